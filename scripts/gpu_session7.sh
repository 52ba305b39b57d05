#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/s7_pytest.log; cat gpurun_out/s7_pytest.log
python scripts/kernel_bench.py --iters 60 --only colorlut64,hsv > gpurun_out/s7_kernel_bench.jsonl 2> gpurun_out/s7_kernel_bench.err
cat gpurun_out/s7_kernel_bench.jsonl | cut -c1-230; tail -3 gpurun_out/s7_kernel_bench.err
python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/s7_bench_n1.json 2> gpurun_out/s7_bench.err
python -c "
import json; d=json.load(open('gpurun_out/s7_bench_n1.json')); print(d['value'], d['roofline']['frac'], d['roofline']['by_content'], d['e2e']['value'])"
tail -3 gpurun_out/s7_bench.err
