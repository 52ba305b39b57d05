#!/bin/bash
# 8-GPU: BASELINE config 5 (8K row-tiled + all-gather) and the bench at N=8 with round-1 final defaults
mkdir -p gpurun_out/s17
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 \
  scripts/config_bench.py --config5 > gpurun_out/s17/config5.jsonl 2> gpurun_out/s17/config5.err
echo "config5 rc=$?"; cat gpurun_out/s17/config5.jsonl
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 \
  bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/s17/bench_n8.json 2> gpurun_out/s17/bench_n8.err
echo "bench rc=$?"; cat gpurun_out/s17/bench_n8.json
