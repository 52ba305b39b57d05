#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/s6_pytest.log; cat gpurun_out/s6_pytest.log
python scripts/kernel_bench.py --iters 60 --only hsv,e2e > gpurun_out/s6_kernel_bench.jsonl 2> gpurun_out/s6_kernel_bench.err
cat gpurun_out/s6_kernel_bench.jsonl | cut -c1-230; tail -3 gpurun_out/s6_kernel_bench.err
python bench.py --steps 20 --warmup 5 > gpurun_out/s6_bench_n1.json 2> gpurun_out/s6_bench.err
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/s6_bench_ref.json 2>> gpurun_out/s6_bench.err
cat gpurun_out/s6_bench_n1.json gpurun_out/s6_bench_ref.json; tail -3 gpurun_out/s6_bench.err
