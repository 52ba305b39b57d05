#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "blockhash or colordetect or videocompare or hash or chain or smoke" > gpurun_out/s9_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s9_pytest.log
python scripts/kernel_bench.py --only videofx,colordetect > gpurun_out/s9_kernel.jsonl 2> gpurun_out/s9_kernel.err
tail -4 gpurun_out/s9_pytest.log; cut -c1-230 gpurun_out/s9_kernel.jsonl
