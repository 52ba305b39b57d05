#!/bin/bash
# session 2: validate the TMA streaming kernel (tests + sanitizer), re-measure, warm-cache ncu, CPU thread scaling
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/s2_pytest.log
cat gpurun_out/s2_pytest.log
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "strides_and_alignment or appendix or lifecycle" > gpurun_out/s2_memcheck.log 2>&1
tail -4 gpurun_out/s2_memcheck.log
timeout 600 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "strides_and_alignment" > gpurun_out/s2_racecheck.log 2>&1
tail -4 gpurun_out/s2_racecheck.log
python scripts/kernel_bench.py --iters 60 --only colorlut,hsv,videofx,e2e > gpurun_out/s2_kernel_bench.jsonl 2> gpurun_out/s2_kernel_bench.err
B200VFX_NO_TMA=1 python scripts/kernel_bench.py --iters 60 --only colorlut > gpurun_out/s2_kernel_bench_notma.jsonl 2>> gpurun_out/s2_kernel_bench.err
python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/s2_bench_n1.json 2> gpurun_out/s2_bench.err
for c in ramps noise natural; do
  ncu --set full --clock-control none --cache-control none --import-source on -k regex:colorlut_memo_stream -s 3 -c 1 -f -o gpurun_out/s2_stream_warm_$c \
      python scripts/ncu_target.py --kernel memo --content $c --launches 8 > gpurun_out/s2_ncu_$c.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:colorlut_memo_stream -s 3 -c 1 -f -o gpurun_out/s2_stream_cold_ramps \
    python scripts/ncu_target.py --kernel memo --content ramps --launches 8 >> gpurun_out/s2_ncu_ramps.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:blockhash -s 2 -c 1 -f -o gpurun_out/s2_blockhash \
    python scripts/ncu_target.py --kernel blockhash --content noise > gpurun_out/s2_ncu_bh.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:hsvfilter -s 2 -c 1 -f -o gpurun_out/s2_hsvfilter \
    python scripts/ncu_target.py --kernel hsvfilter --content noise >> gpurun_out/s2_ncu_bh.log 2>&1
# CPU oracle thread scaling + cgroup limits
(cat /sys/fs/cgroup/cpu.max; nproc; lscpu | grep -E 'Model name|Socket|Core|Thread|MHz' ) > gpurun_out/s2_cpuinfo.txt 2>&1
python - > gpurun_out/s2_cpu_scaling.txt 2>&1 <<'EOF'
import sys, time, os
sys.path.insert(0, "gst-plugin-rs_b200"); sys.path.insert(0, "tests")
import numpy as np, oracle_binding as orc
from b200vfx import synth
cube = orc.cube_parse(synth.cube_text_3d(33, "mix"))
f = synth.frame_noise("RGBA", 3840, 2160, 1)
out = np.zeros_like(f)
for t in (1, 2, 4, 8, 16, 32, 64, 128):
    orc.colorlut_apply(cube, "RGBA", 3840, 2160, f, threads=t, out=out)
    t0 = time.perf_counter(); n = 3 if t < 8 else 10
    for _ in range(n): orc.colorlut_apply(cube, "RGBA", 3840, 2160, f, threads=t, out=out)
    print(t, "threads", round(n / (time.perf_counter() - t0), 2), "fps")
EOF
cat gpurun_out/s2_cpu_scaling.txt gpurun_out/s2_cpuinfo.txt
cat gpurun_out/s2_kernel_bench.jsonl | head -60
cat gpurun_out/s2_bench_n1.json
tail -3 gpurun_out/s2_bench.err gpurun_out/s2_kernel_bench.err
