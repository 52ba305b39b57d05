#!/bin/bash
# 2 GPUs: protocol tests of the fused tile-gather kernel, then the torchrun parity + timing check
mkdir -p gpurun_out/s18
timeout 300 python -m pytest tests/test_gpu_tilegather.py -x -q 2>&1 | tail -15
N=${1:-2}
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
  scripts/tile_gather_check.py --small > gpurun_out/s18/small_n$N.txt 2> gpurun_out/s18/small_n$N.err
echo "small rc=$?"; cat gpurun_out/s18/small_n$N.txt; tail -5 gpurun_out/s18/small_n$N.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 \
  scripts/tile_gather_check.py > gpurun_out/s18/full_n$N.txt 2> gpurun_out/s18/full_n$N.err
echo "full rc=$?"; cat gpurun_out/s18/full_n$N.txt; tail -5 gpurun_out/s18/full_n$N.err
