#!/bin/bash
# N GPUs: fused tile gather variants incl. NVSwitch multicast + copy-engine yardstick
mkdir -p gpurun_out
N=${1:-2}
NCCL_DEBUG=WARN timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 scripts/tile_gather_check.py --iters 60 > gpurun_out/s16_tile_gather_n$N.txt 2>&1
echo "rc=$?"
grep -v "^rank [1-9]" gpurun_out/s16_tile_gather_n$N.txt | tail -12 | cut -c1-1800
