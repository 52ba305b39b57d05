#!/bin/bash
mkdir -p gpurun_out/prof21
O=gpurun_out/prof21
NCU="ncu --set full --clock-control none --import-source on -f"
$NCU -k regex:luma_ -s 4 -c 2 -o $O/hash_resize python scripts/ncu_target.py --kernel hash --content noise --launches 4 > $O/ncu.log 2>&1
$NCU -k regex:blockhash_rows -s 2 -c 1 -o $O/blockhash_rows python scripts/ncu_target.py --kernel blockhash --content noise --launches 4 >> $O/ncu.log 2>&1
for r in $O/*.ncu-rep; do ncu -i $r --page raw --csv > ${r%.ncu-rep}.raw.csv 2>/dev/null; done
rm -f $O/*.ncu-rep
tail -3 $O/ncu.log; ls -la $O
