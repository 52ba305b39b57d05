#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "stress or memo_tile" > gpurun_out/s11_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s11_pytest.log
python scripts/exp_memo_tile.py > gpurun_out/s11_memo_tile.jsonl 2> gpurun_out/s11_memo_tile.err
for tool in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 python scripts/racecheck_target.py > gpurun_out/s11_sanitizer_$tool.txt 2>&1
  echo "$tool rc=$?" >> gpurun_out/s11_sanitizer_$tool.txt
done
tail -3 gpurun_out/s11_pytest.log; cat gpurun_out/s11_memo_tile.jsonl; for t in memcheck racecheck synccheck; do tail -4 gpurun_out/s11_sanitizer_$t.txt; done
