#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -4
timeout 600 python scripts/config_bench.py > gpurun_out/s28_configs.jsonl 2> gpurun_out/s28_configs.err; cut -c1-420 gpurun_out/s28_configs.jsonl; tail -2 gpurun_out/s28_configs.err
