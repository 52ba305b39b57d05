#!/bin/bash
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 6 --warmup 3 > gpurun_out/s8_bench_n2.json 2> gpurun_out/s8_bench_n2.err
echo "rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/s8_bench_n2.json'))
print(d['value'], d['roofline']['frac'])
print(json.dumps(d['roofline'].get('multi_gpu'), indent=1))
e=d['e2e']; print(e['value'], e.get('host_ceiling'))
print(d['config'])
PY
python bench.py --impl reference --gpus 2 --steps 3 --warmup 1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['config'])"
tail -5 gpurun_out/s8_bench_n2.err
