#!/bin/bash
# colordetect histogram kernel + 3-bpp memo kernels: parity, then per-kernel timings
mkdir -p gpurun_out/s21
timeout 600 python -m pytest tests/test_colordetect.py tests/test_gpu_parity.py -m gpu -x -q -k "colordetect or hsv" > gpurun_out/s21/pytest.txt 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/s21/pytest.txt
timeout 300 python scripts/kernel_bench.py --only colordetect,hsv24 --iters 40 > gpurun_out/s21/kernel_bench.jsonl 2> gpurun_out/s21/kernel_bench.err; echo "kb rc=$?"
cat gpurun_out/s21/kernel_bench.jsonl; tail -3 gpurun_out/s21/kernel_bench.err
