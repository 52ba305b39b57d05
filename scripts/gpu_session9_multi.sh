#!/bin/bash
# multi-GPU validation: bench under torchrun at N=2 (own arm + reference arm), plus a 2-rank NCCL tile all-gather parity check
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/s9_bench_n2.json 2> gpurun_out/s9_bench_n2.err
cat gpurun_out/s9_bench_n2.json | cut -c1-3000; tail -5 gpurun_out/s9_bench_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 5 --warmup 3 > gpurun_out/s9_bench_ref_n2.json 2>> gpurun_out/s9_bench_n2.err
cat gpurun_out/s9_bench_ref_n2.json | cut -c1-400
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 scripts/nccl_tiles_check.py > gpurun_out/s9_nccl_tiles.txt 2>&1
tail -5 gpurun_out/s9_nccl_tiles.txt
python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/s9_bench_n1.json 2>> gpurun_out/s9_bench_n2.err
python -c "
import json
for f in ('gpurun_out/s9_bench_n1.json','gpurun_out/s9_bench_n2.json'):
    d=json.load(open(f)); print(f, d['n_gpus'], round(d['value']), round(d['e2e']['value']), d.get('allgather'))"
