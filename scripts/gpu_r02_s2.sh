#!/bin/bash
# round-2 GPU session 2: full GPU test suite + hsv kernel timings
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/s2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s2_pytest.log
python scripts/kernel_bench.py --only hsv,hsvanim,hsv24 > gpurun_out/s2_kernel_hsv.jsonl 2> gpurun_out/s2_kernel_hsv.err
tail -3 gpurun_out/s2_pytest.log
