#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_variants.py tests/test_gpu_parity.py tests/test_videocompare_hashes.py tests/test_elements.py -q -m gpu -x -k "blockhash or videocompare or hash or pdl" 2>&1 | tail -4
timeout 600 python scripts/exp_blockhash.py 2>&1 | tee gpurun_out/s17_blockhash.jsonl | cut -c1-400
