#!/bin/bash
mkdir -p gpurun_out

timeout 600 python scripts/exp_colordetect_split.py 2>&1 | tee gpurun_out/s30_cd_split.jsonl | cut -c1-300
