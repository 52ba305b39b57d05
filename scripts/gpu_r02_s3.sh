#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "memo_tile or persistent_lookup" > gpurun_out/s3_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s3_pytest.log
python scripts/exp_memo_tile.py > gpurun_out/s3_memo_tile.jsonl 2> gpurun_out/s3_memo_tile.err
ncu --set full --clock-control none --import-source on -k regex:hsv_direct_map -s 2 -c 1 -o gpurun_out/s3_hsvdirect -f python scripts/ncu_target.py --kernel hsvfilter --content ramps --opt hsv_memo=0 --launches 4 > gpurun_out/s3_ncu.log 2>&1
tail -3 gpurun_out/s3_pytest.log; cat gpurun_out/s3_memo_tile.jsonl
