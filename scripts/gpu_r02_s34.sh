#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_variants.py -q -m gpu -x -k "async or mixed or chain or order" 2>&1 | tail -4
