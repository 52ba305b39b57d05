#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "rc=$?"
python - $N <<'PY'
import json, sys
n = sys.argv[1]
d = json.load(open('gpurun_out/bench_n%s.json' % n))
print(d['value'], d['roofline']['frac'], d['n_gpus'])
print(json.dumps(d['roofline'].get('multi_gpu'), indent=1))
e = d['e2e']; print(e['value'], e.get('host_ceiling'))
print(d.get('allgather'))
PY
tail -3 gpurun_out/bench_n$N.err
