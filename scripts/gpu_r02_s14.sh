#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "hsv" > gpurun_out/s14_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s14_pytest.log
python scripts/kernel_bench.py --only hsvanim,colorlut64 > gpurun_out/s14_kernel.jsonl 2> gpurun_out/s14_kernel.err
ncu --set full --clock-control none -k regex:hsv_direct_map -s 2 -c 1 -o gpurun_out/s14_hsv -f python scripts/ncu_target.py --kernel hsvfilter --content ramps --opt hsv_memo=0 --launches 4 > gpurun_out/s14_ncu.log 2>&1
ncu -i gpurun_out/s14_hsv.ncu-rep --page raw --csv > gpurun_out/s14_hsv.raw.csv 2>/dev/null; rm -f gpurun_out/s14_hsv.ncu-rep
tail -4 gpurun_out/s14_pytest.log; cut -c1-200 gpurun_out/s14_kernel.jsonl
