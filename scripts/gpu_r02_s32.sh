#!/bin/bash
timeout 600 python -m pytest tests/test_golden_r02.py tests/test_golden.py -q -m gpu 2>&1 | tail -5
