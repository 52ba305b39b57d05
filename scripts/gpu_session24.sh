#!/bin/bash
mkdir -p gpurun_out/s24
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s24/pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/s24/pytest_gpu.txt
timeout 300 python scripts/kernel_bench.py --only colordetect,videofx --iters 40 > gpurun_out/s24/kernel_bench.jsonl 2> gpurun_out/s24/kernel_bench.err; echo "kb rc=$?"
cut -c1-210 gpurun_out/s24/kernel_bench.jsonl; tail -3 gpurun_out/s24/kernel_bench.err
python scripts/cd_probe.py 2>&1 | tail -8
