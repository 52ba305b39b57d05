#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/s3_pytest.log; cat gpurun_out/s3_pytest.log
python scripts/sweep_memo.py > gpurun_out/s3_sweep.jsonl 2> gpurun_out/s3_sweep.err
python scripts/kernel_bench.py --iters 60 --only hsv,videofx > gpurun_out/s3_kernel_bench.jsonl 2> gpurun_out/s3_kernel_bench.err
cat gpurun_out/s3_sweep.jsonl; tail -2 gpurun_out/s3_sweep.err; cat gpurun_out/s3_kernel_bench.jsonl; tail -2 gpurun_out/s3_kernel_bench.err
