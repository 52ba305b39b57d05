#!/bin/bash
# N GPUs: torchrun parity + timing of the fused tile-gather kernel at BASELINE config 5 shape
N=${1:-8}
mkdir -p gpurun_out/s19
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 \
  scripts/tile_gather_check.py > gpurun_out/s19/full_n$N.txt 2> gpurun_out/s19/full_n$N.err
echo "full rc=$?"; cat gpurun_out/s19/full_n$N.txt; grep -v "^\*\*\*\|OMP_NUM\|^$" gpurun_out/s19/full_n$N.err | tail -5
