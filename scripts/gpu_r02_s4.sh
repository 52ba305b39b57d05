#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "hsv" > gpurun_out/s4_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s4_pytest.log
python scripts/kernel_bench.py --only hsvanim,hsv > gpurun_out/s4_kernel_hsv.jsonl 2> gpurun_out/s4_kernel_hsv.err
tail -3 gpurun_out/s4_pytest.log; grep -E "animated|direct" gpurun_out/s4_kernel_hsv.jsonl | cut -c1-160
