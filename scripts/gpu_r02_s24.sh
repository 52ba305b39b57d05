#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_round2.py -q -m gpu -x -k "async" 2>&1 | tail -6
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/s24_bench.json 2> gpurun_out/s24_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open('gpurun_out/s24_bench.json'))
e = d['e2e']
print(d['value'], d['roofline']['frac'])
print({k: e[k] for k in ('value', 'ms_per_frame', 'mode', 'synchronous_calls', 'async_error', 'pcie_concurrent_memcpy')})
PY
tail -3 gpurun_out/s24_bench.err
