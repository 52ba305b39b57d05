#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "colorlut or smoke or golden" > gpurun_out/s10_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s10_pytest.log
python scripts/kernel_bench.py --only colorlut64,videofx > gpurun_out/s10_kernel.jsonl 2> gpurun_out/s10_kernel.err
python scripts/kernel_bench.py --only colorlut --iters 30 2>/dev/null | grep -E "direct" > gpurun_out/s10_kernel_direct8.jsonl
tail -4 gpurun_out/s10_pytest.log; cut -c1-250 gpurun_out/s10_kernel.jsonl; cut -c1-200 gpurun_out/s10_kernel_direct8.jsonl; tail -3 gpurun_out/s10_kernel.err
