#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/s12_pytest.log; cat gpurun_out/s12_pytest.log
python scripts/sweep_memo.py > gpurun_out/s12_sweep.jsonl 2> gpurun_out/s12_sweep.err
cat gpurun_out/s12_sweep.jsonl | cut -c1-260; tail -2 gpurun_out/s12_sweep.err
python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/s12_bench_n1.json 2> gpurun_out/s12_bench.err
python -c "
import json; d=json.load(open('gpurun_out/s12_bench_n1.json')); print(d['value'], d['roofline']['frac'], d['roofline']['by_content'], d['e2e']['value'], d['roofline']['traffic'])"
tail -2 gpurun_out/s12_bench.err
ncu --set full --clock-control none --cache-control none --import-source on -k regex:colorlut_memo_apply -s 3 -c 1 -f -o gpurun_out/s12_noise_warm_l2persist \
      python scripts/ncu_target.py --kernel memo --content noise --launches 8 > gpurun_out/s12_ncu.log 2>&1
