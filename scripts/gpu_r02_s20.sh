#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_videocompare_hashes.py tests/test_elements.py tests/test_gpu_round2.py -q -m gpu -x 2>&1 | tail -4
timeout 600 python scripts/kernel_bench.py --only hashes 2>&1 | tee gpurun_out/s20_hashes.jsonl | cut -c1-330
