#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py smoke 2>&1 | tail -2
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python scripts/config_bench.py > gpurun_out/s13_configs.jsonl 2> gpurun_out/s13_configs.err
cat gpurun_out/s13_configs.jsonl; tail -3 gpurun_out/s13_configs.err
python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/s13_bench_n1.json 2> gpurun_out/s13_bench.err
python -c "
import json; d=json.load(open('gpurun_out/s13_bench_n1.json')); print(d['value'], d['roofline']['frac'], d['roofline']['by_content'], d['e2e']['value'], d['roofline']['traffic'])"
tail -2 gpurun_out/s13_bench.err
