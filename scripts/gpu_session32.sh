#!/bin/bash
mkdir -p gpurun_out/s32
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s32/pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/s32/pytest_gpu.txt
python scripts/exp_memo_ctas.py 2>&1 | tail -4
timeout 300 python scripts/kernel_bench.py --only hsv,hsv24 --iters 60 2>/dev/null | grep -E "memo" | cut -c1-170
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/s32/bench_n1.json 2> gpurun_out/s32/bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/s32/bench_n1.json"))
print("value", round(d["value"]), "frac", round(d["roofline"]["frac"],3), {k: round(v["us_per_frame"],2) for k,v in d["roofline"]["by_content"].items()}, "e2e", round(d["e2e"]["value"],1))
PY
