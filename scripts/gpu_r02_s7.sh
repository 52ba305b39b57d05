#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "convert or chain or persistent or 1d_device" > gpurun_out/s7_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s7_pytest.log
python scripts/kernel_bench.py --only hsvanim > gpurun_out/s7_kernel_hsv.jsonl 2> gpurun_out/s7_kernel_hsv.err
python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/s7_bench.json 2> gpurun_out/s7_bench.err
tail -5 gpurun_out/s7_pytest.log; cat gpurun_out/s7_kernel_hsv.jsonl | cut -c1-150
python -c "
import json
d=json.load(open('gpurun_out/s7_bench.json'))
print(d['value'], d['roofline']['frac'], {k:(round(v['us_per_frame'],2), round(v['frac'],3)) for k,v in d['roofline']['by_content'].items()}, d['e2e']['value'])
"
