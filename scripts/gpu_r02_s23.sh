#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_convert.py -q -m gpu -x 2>&1 | tail -6
timeout 600 python scripts/kernel_bench.py --only planar 2>&1 | tee gpurun_out/s23_planar.jsonl | cut -c1-260
