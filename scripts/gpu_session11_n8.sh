#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/s11_bench_n8.json 2> gpurun_out/s11_bench_n8.err
tail -c 2500 gpurun_out/s11_bench_n8.json; echo; grep -v "^\*\*\*\|OMP_NUM\|^$" gpurun_out/s11_bench_n8.err | tail -5
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 scripts/nccl_tiles_check.py > gpurun_out/s11_nccl_tiles_n8.txt 2>&1
grep "^rank" gpurun_out/s11_nccl_tiles_n8.txt | sort | head -8
python -m pytest tests/test_gpu_variants.py -q -m gpu -k "every_visible_device" 2>&1 | tail -2
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 4 --steps 20 --warmup 5 --no-e2e > gpurun_out/s11_bench_n4.json 2>> gpurun_out/s11_bench_n8.err
python -c "
import json
for f in ('gpurun_out/s11_bench_n4.json','gpurun_out/s11_bench_n8.json'):
    d=json.loads([l for l in open(f) if l.startswith('{')][-1]); print(f, d['n_gpus'], round(d['value']), d.get('e2e',{}).get('value'), d.get('allgather'))"
