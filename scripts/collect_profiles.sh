#!/bin/bash
# profile collection: everything that profiles/r02_* is digested from (run on one B200 through gpurun)
mkdir -p gpurun_out/prof
O=gpurun_out/prof
python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
python bench.py --impl reference --steps 20 --warmup 3 > $O/bench_ref_n1.json 2> $O/bench_ref_n1.err
sleep 3; python bench.py > $O/bench_n1.json 2> $O/bench_n1.err
sleep 3; python bench.py --steps 20 --warmup 5 > $O/bench_n1_run2.json 2> $O/bench_n1_run2.err
python __graft_entry__.py smoke > $O/smoke.log 2>&1; tail -1 $O/smoke.log
python scripts/kernel_bench.py > $O/kernel_matrix.jsonl 2> $O/kernel_matrix.err
python scripts/config_bench.py > $O/configs.jsonl 2> $O/configs.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 80 -c 400 --csv --log-file $O/launches_bench.csv python bench.py --steps 2 --warmup 3 --gop 8 --no-cpu --no-e2e > $O/launches_bench.log 2>&1
NCU="ncu --set full --clock-control none --import-source on -f"
$NCU -k regex:colorlut_memo_apply -s 2 -c 1 -o $O/memo_ramps python scripts/ncu_target.py --kernel memo --content ramps --launches 4 > $O/ncu.log 2>&1
$NCU -k regex:colorlut_memo_apply -s 2 -c 1 -o $O/memo_noise python scripts/ncu_target.py --kernel memo --content noise --launches 4 >> $O/ncu.log 2>&1
$NCU -k regex:colorlut_memo_apply -s 2 -c 1 -o $O/memo_natural python scripts/ncu_target.py --kernel memo --content natural --launches 4 >> $O/ncu.log 2>&1
$NCU -k regex:colorlut_direct64x4 -s 2 -c 1 -o $O/rgba64_ramps python scripts/ncu_target.py --kernel direct64 --content ramps --launches 4 >> $O/ncu.log 2>&1
$NCU -k regex:colorlut_direct64x4 -s 2 -c 1 -o $O/rgba64_noise python scripts/ncu_target.py --kernel direct64 --content noise --launches 4 >> $O/ncu.log 2>&1
$NCU -k regex:hsv_direct_map -s 2 -c 1 -o $O/hsvfilter_direct python scripts/ncu_target.py --kernel hsvfilter --content ramps --opt hsv_memo=0 --launches 4 >> $O/ncu.log 2>&1
$NCU -k regex:blockhash_rows -s 2 -c 1 -o $O/blockhash python scripts/ncu_target.py --kernel blockhash --content noise --launches 4 >> $O/ncu.log 2>&1
$NCU -k regex:luma_ -s 4 -c 2 -o $O/hash_resize python scripts/ncu_target.py --kernel hash --content noise --launches 4 >> $O/ncu.log 2>&1
$NCU -k regex:map_u32_kernel -s 2 -c 1 -o $O/colorlut_fmt python scripts/ncu_target.py --kernel fmt --content ramps --launches 4 >> $O/ncu.log 2>&1
$NCU -k regex:colorlut_i420 -s 2 -c 1 -o $O/colorlut_i420 python scripts/ncu_target.py --kernel planar --content natural --launches 4 >> $O/ncu.log 2>&1
$NCU -k regex:colordetect_hist -s 2 -c 1 -o $O/colordetect_q10 python scripts/ncu_target.py --kernel colordetect --quality 10 --content noise --launches 4 >> $O/ncu.log 2>&1
# the reports are ~12 MB each and gpurun brings back at most 64 MiB: export the raw metric pages here, keep two reports
for r in $O/*.ncu-rep; do ncu -i $r --page raw --csv > ${r%.ncu-rep}.raw.csv 2>/dev/null; done
rm -f $O/*.ncu-rep
tail -3 $O/pytest_gpu.log; ls -la $O | head -40; du -sh gpurun_out
