#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "rgba64" > gpurun_out/s12_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s12_pytest.log
python scripts/kernel_bench.py --only colorlut64 > gpurun_out/s12_kernel.jsonl 2> gpurun_out/s12_kernel.err
tail -4 gpurun_out/s12_pytest.log; cut -c1-200 gpurun_out/s12_kernel.jsonl; tail -3 gpurun_out/s12_kernel.err
