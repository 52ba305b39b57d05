#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -q -m gpu -x -k "hsv" 2>&1 | tail -3
timeout 300 python scripts/kernel_bench.py --only hsvanim 2>&1 | grep hsvfilter | cut -c1-200 | tee gpurun_out/s33_hsvanim.jsonl
