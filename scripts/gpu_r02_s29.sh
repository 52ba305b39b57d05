#!/bin/bash
mkdir -p gpurun_out
python scripts/pcie_peak.py 2>/dev/null | tail -3 | cut -c1-300
timeout 600 python scripts/config_bench.py > gpurun_out/s29_configs.jsonl 2> gpurun_out/s29_configs.err; cut -c1-420 gpurun_out/s29_configs.jsonl; tail -2 gpurun_out/s29_configs.err
