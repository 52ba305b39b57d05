#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/s4_pytest.log; cat gpurun_out/s4_pytest.log
python scripts/sweep_memo.py > gpurun_out/s4_sweep.jsonl 2> gpurun_out/s4_sweep.err
python scripts/kernel_bench.py --iters 60 --only colorlut,colorlut64 > gpurun_out/s4_kernel_bench.jsonl 2> gpurun_out/s4_kernel_bench.err
python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/s4_bench_n1.json 2> gpurun_out/s4_bench.err
ncu --set full --clock-control none --import-source on -k regex:colorlut_direct -s 2 -c 1 -f -o gpurun_out/s4_direct64_ramps \
    python scripts/ncu_target.py --kernel direct64 --content ramps > gpurun_out/s4_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:colorlut_direct -s 2 -c 1 -f -o gpurun_out/s4_direct64_noise \
    python scripts/ncu_target.py --kernel direct64 --content noise >> gpurun_out/s4_ncu.log 2>&1
cat gpurun_out/s4_sweep.jsonl; tail -2 gpurun_out/s4_sweep.err; cat gpurun_out/s4_kernel_bench.jsonl | cut -c1-180; tail -2 gpurun_out/s4_kernel_bench.err; cat gpurun_out/s4_bench_n1.json; tail -2 gpurun_out/s4_bench.err
