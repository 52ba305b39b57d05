#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/s15_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/s15_smoke.log
python scripts/exp_two_tables.py > gpurun_out/s15_two_tables.jsonl 2> gpurun_out/s15_two_tables.err
python -m pytest tests -m gpu -q > gpurun_out/s15_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s15_pytest.log
for i in 1 2 3; do python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bench run', d['value'], d['roofline']['frac'])"; done > gpurun_out/s15_bench3.txt
tail -2 gpurun_out/s15_smoke.log; cat gpurun_out/s15_two_tables.jsonl; tail -3 gpurun_out/s15_pytest.log; cat gpurun_out/s15_bench3.txt
