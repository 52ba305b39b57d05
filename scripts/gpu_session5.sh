#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/s5_pytest.log; cat gpurun_out/s5_pytest.log
python scripts/kernel_bench.py --iters 60 --only colorlut64,hsv,videofx,e2e > gpurun_out/s5_kernel_bench.jsonl 2> gpurun_out/s5_kernel_bench.err
cat gpurun_out/s5_kernel_bench.jsonl | cut -c1-220; tail -3 gpurun_out/s5_kernel_bench.err
ncu --set full --clock-control none --import-source on -k regex:colorlut_direct -s 2 -c 1 -f -o gpurun_out/s5_direct64_ramps \
    python scripts/ncu_target.py --kernel direct64 --content ramps > gpurun_out/s5_ncu.log 2>&1
