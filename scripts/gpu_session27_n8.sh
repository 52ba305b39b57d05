#!/bin/bash
# 8 GPUs: fused tile gather parity + timing at BASELINE config 5 (8K frame, 65^3), then the bench line with the reassembly legs
N=${1:-8}
mkdir -p gpurun_out/s27
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29552 \
  scripts/tile_gather_check.py --iters 60 > gpurun_out/s27/full_n$N.txt 2> gpurun_out/s27/full_n$N.err
echo "full rc=$?"; grep -v "^\*\*\*\|OMP_NUM\|^$" gpurun_out/s27/full_n$N.txt | cut -c1-1500; tail -3 gpurun_out/s27/full_n$N.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29553 \
  bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/s27/bench_n$N.json 2> gpurun_out/s27/bench_n$N.err
echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/s27/bench_n$N.json')); print(d['value'], d['e2e']['value'], d['allgather'])"
