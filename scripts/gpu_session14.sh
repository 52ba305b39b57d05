#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python scripts/e2e_sweep.py > gpurun_out/s14_e2e.jsonl 2> gpurun_out/s14_e2e.err
cat gpurun_out/s14_e2e.jsonl; tail -2 gpurun_out/s14_e2e.err
python scripts/config_bench.py > gpurun_out/s14_configs.jsonl 2> gpurun_out/s14_configs.err
cat gpurun_out/s14_configs.jsonl; tail -3 gpurun_out/s14_configs.err
