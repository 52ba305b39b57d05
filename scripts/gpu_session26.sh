#!/bin/bash
mkdir -p gpurun_out/s26
timeout 900 python -m pytest tests -m gpu -x -q -k "colorlut or golden or smoke or tile" > gpurun_out/s26/pytest.txt 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/s26/pytest.txt
timeout 300 python scripts/kernel_bench.py --only colorlut64 --iters 60 > gpurun_out/s26/kb64.jsonl 2> gpurun_out/s26/kb.err; cut -c1-200 gpurun_out/s26/kb64.jsonl
timeout 300 python scripts/kernel_bench.py --only colorlut --iters 40 2>> gpurun_out/s26/kb.err | grep -E "direct|memo" | grep '"lut": 33' | cut -c1-200
