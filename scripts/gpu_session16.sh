python -m pytest tests -m gpu -x -q 2>&1 | tail -3
#!/bin/bash
# profile artefacts for profiles/: launch list of the bench command + full captures of the dominant kernels
mkdir -p gpurun_out/p
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/p/launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --gop 8 --no-cpu --no-e2e > gpurun_out/p/bench_under_ncu.log 2>&1
for c in ramps noise natural; do
  ncu --set full --clock-control none --import-source on -k regex:colorlut_memo_apply -s 3 -c 1 -f -o gpurun_out/p/memo_apply_${c}_cold \
      python scripts/ncu_target.py --kernel memo --content $c --launches 8 > gpurun_out/p/ncu_memo_$c.log 2>&1
  ncu --set full --clock-control none --cache-control none --import-source on -k regex:colorlut_memo_apply -s 3 -c 1 -f -o gpurun_out/p/memo_apply_${c}_warm \
      python scripts/ncu_target.py --kernel memo --content $c --launches 8 >> gpurun_out/p/ncu_memo_$c.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:colorlut_direct -s 2 -c 1 -f -o gpurun_out/p/direct64_ramps \
    python scripts/ncu_target.py --kernel direct64 --content ramps > gpurun_out/p/ncu_misc.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:map_u32 -s 3 -c 1 -f -o gpurun_out/p/hsvfilter_memo_ramps \
    python scripts/ncu_target.py --kernel hsvfilter --content ramps --launches 16 >> gpurun_out/p/ncu_misc.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:hsvfilter_kernel -s 1 -c 1 -f -o gpurun_out/p/hsvfilter_direct_noise \
    python scripts/ncu_target.py --kernel hsvfilter --content noise --launches 2 >> gpurun_out/p/ncu_misc.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:blockhash -s 3 -c 1 -f -o gpurun_out/p/blockhash_noise \
    python scripts/ncu_target.py --kernel blockhash --content noise >> gpurun_out/p/ncu_misc.log 2>&1
python scripts/kernel_bench.py --iters 60 > gpurun_out/p/kernel_bench.jsonl 2> gpurun_out/p/kernel_bench.err
python scripts/sweep_memo.py > gpurun_out/p/sweep_memo.jsonl 2>> gpurun_out/p/kernel_bench.err
python bench.py --steps 30 --warmup 5 > gpurun_out/p/bench_n1.json 2> gpurun_out/p/bench_n1.err
python bench.py --impl reference --steps 30 --warmup 3 > gpurun_out/p/bench_ref_n1.json 2>> gpurun_out/p/bench_n1.err
ls -la gpurun_out/p | head -40
tail -2 gpurun_out/p/kernel_bench.err gpurun_out/p/bench_n1.err
cat gpurun_out/p/bench_n1.json | cut -c1-600
