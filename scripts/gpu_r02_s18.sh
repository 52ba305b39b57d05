#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/exp_blockhash.py 2>&1 | tee gpurun_out/s18_blockhash.jsonl | cut -c1-500
