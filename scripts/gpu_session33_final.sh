#!/bin/bash
# final evidence after the persistent-grid change: tests, smoke, bench, launch list, full captures of the dominant kernel
mkdir -p gpurun_out/s33
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s33/pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/s33/pytest_gpu.txt
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/s33/bench_n1.json 2> gpurun_out/s33/bench_n1.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/s33/bench_ref_n1.json 2>> gpurun_out/s33/bench_n1.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/s33/bench_n1.json"))
print("value", round(d["value"]), "frac", round(d["roofline"]["frac"],3), {k: round(v["us_per_frame"],2) for k,v in d["roofline"]["by_content"].items()}, "e2e", round(d["e2e"]["value"],1), "launches", d["gpu_launches"])
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/s33/launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --gop 8 --no-cpu --no-e2e > gpurun_out/s33/bench_under_ncu.log 2>&1
for c in ramps noise natural; do
  ncu --set full --clock-control none --import-source on -k regex:colorlut_memo_apply -s 3 -c 1 -f -o gpurun_out/s33/memo_apply_${c}_cold \
      python scripts/ncu_target.py --kernel memo --content $c --launches 8 > gpurun_out/s33/ncu_memo_$c.log 2>&1
done
ls gpurun_out/s33
