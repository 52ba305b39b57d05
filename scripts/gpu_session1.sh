#!/bin/bash
# one gpurun call: matrix of kernel timings, the bench line, ncu launch list and full captures
mkdir -p gpurun_out
python scripts/kernel_bench.py --iters 60 > gpurun_out/kernel_bench.jsonl 2> gpurun_out/kernel_bench.err
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
python bench.py --steps 20 --warmup 5 --mode 1 --no-cpu --no-e2e > gpurun_out/bench_n1_direct.json 2>> gpurun_out/bench_n1.err
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref.json 2>> gpurun_out/bench_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --gop 8 --no-cpu --no-e2e > gpurun_out/bench_under_ncu.log 2>&1
for c in ramps noise natural; do
  ncu --set full --clock-control none --import-source on -k regex:colorlut_memo_apply -s 2 -c 2 -f -o gpurun_out/memo_$c \
      python scripts/ncu_target.py --kernel memo --content $c > gpurun_out/ncu_memo_$c.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:colorlut_direct -s 2 -c 1 -f -o gpurun_out/direct_ramps \
    python scripts/ncu_target.py --kernel direct --content ramps > gpurun_out/ncu_direct.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:colorlut_direct -s 2 -c 1 -f -o gpurun_out/direct_noise \
    python scripts/ncu_target.py --kernel direct --content noise >> gpurun_out/ncu_direct.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:hsvfilter -s 2 -c 1 -f -o gpurun_out/hsvfilter_noise \
    python scripts/ncu_target.py --kernel hsvfilter --content noise > gpurun_out/ncu_hsv.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:blockhash -s 2 -c 1 -f -o gpurun_out/blockhash \
    python scripts/ncu_target.py --kernel blockhash --content noise >> gpurun_out/ncu_hsv.log 2>&1
tail -3 gpurun_out/bench_n1.err
cat gpurun_out/bench_n1.json
head -50 gpurun_out/kernel_bench.jsonl
