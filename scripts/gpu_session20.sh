#!/bin/bash
# re-entry validation: full GPU suite, smoke, bench both arms (1 GPU)
mkdir -p gpurun_out/s20
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s20/pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/s20/pytest_gpu.txt
timeout 200 python __graft_entry__.py smoke > gpurun_out/s20/smoke.txt 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/s20/smoke.txt
timeout 400 python bench.py --steps 30 --warmup 5 > gpurun_out/s20/bench_n1.json 2> gpurun_out/s20/bench_n1.err; echo "bench rc=$?"
cut -c1-1500 gpurun_out/s20/bench_n1.json
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/s20/bench_ref_n1.json 2>> gpurun_out/s20/bench_n1.err
cut -c1-400 gpurun_out/s20/bench_ref_n1.json
