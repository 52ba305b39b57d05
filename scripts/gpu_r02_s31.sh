#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
python __graft_entry__.py smoke 2>&1 | tail -1
timeout 300 python scripts/kernel_bench.py --only videofx 2>&1 | grep -E "colordetect|blockhash" | cut -c1-260 | tee gpurun_out/s31_videofx.jsonl
sleep 3
python bench.py > gpurun_out/s31_bench.json 2> gpurun_out/s31_bench.err; python -c "
import json; d=json.load(open('gpurun_out/s31_bench.json')); print(d['value'], d['roofline']['frac'], d['e2e']['value'], d['clocks'])"
