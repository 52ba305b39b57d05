#!/bin/bash
mkdir -p gpurun_out/s23
for c in ramps noise; do
  ncu --set full --clock-control none --import-source on -k regex:colordetect_hist -s 2 -c 1 -f -o gpurun_out/s23/cd_${c}_q1 \
      python scripts/ncu_target.py --kernel colordetect --content $c --launches 4 --quality 1 > gpurun_out/s23/ncu_$c.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:colordetect_hist -s 2 -c 1 -f -o gpurun_out/s23/cd_ramps_q10 \
      python scripts/ncu_target.py --kernel colordetect --content ramps --launches 4 --quality 10 > gpurun_out/s23/ncu_q10.log 2>&1
ls -la gpurun_out/s23
