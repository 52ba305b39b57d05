#!/bin/bash
# final round-1 evidence: smoke, bench (both arms), launch list of the bench command, full kernel matrix
mkdir -p gpurun_out/s30
timeout 300 python __graft_entry__.py smoke > gpurun_out/s30/smoke.txt 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/s30/smoke.txt
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/s30/bench_n1.json 2> gpurun_out/s30/bench_n1.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/s30/bench_ref_n1.json 2>> gpurun_out/s30/bench_n1.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/s30/bench_n1.json"))
print("value", round(d["value"]), "frac", round(d["roofline"]["frac"],3), {k: round(v["us_per_frame"],2) for k,v in d["roofline"]["by_content"].items()})
print("e2e", round(d["e2e"]["value"],1), d["e2e"].get("pcie_concurrent_memcpy"), "cpu", d.get("cpu_baseline",{}).get("value"), "clocks", d["clocks"])
r=json.load(open("gpurun_out/s30/bench_ref_n1.json")); print("ref", r["value"], r["cpu_baseline"]["cores"])
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/s30/launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --gop 8 --no-cpu --no-e2e > gpurun_out/s30/bench_under_ncu.log 2>&1
python scripts/kernel_bench.py --iters 60 > gpurun_out/s30/kernel_bench.jsonl 2> gpurun_out/s30/kernel_bench.err; echo "kb rc=$?"; wc -l gpurun_out/s30/kernel_bench.jsonl
python scripts/config_bench.py > gpurun_out/s30/configs.jsonl 2> gpurun_out/s30/configs.err; echo "cfg rc=$?"; cut -c1-300 gpurun_out/s30/configs.jsonl
