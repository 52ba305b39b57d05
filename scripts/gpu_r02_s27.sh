#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/exp_e2e_async.py 2>&1 | tee gpurun_out/s27_e2e_async.jsonl | cut -c1-300
