#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_round2.py -q -m gpu -x -k "async" 2>&1 | grep -E "Error|error|passed|failed" | head -12
timeout 600 python scripts/exp_ab_drift.py 2>&1 | tee gpurun_out/s26_ab_drift.jsonl | cut -c1-1500
