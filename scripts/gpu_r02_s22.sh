#!/bin/bash
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python scripts/racecheck_target.py > gpurun_out/s22_sanitizer_$tool.txt 2>&1
  echo "$tool rc=$?"; tail -3 gpurun_out/s22_sanitizer_$tool.txt
done
