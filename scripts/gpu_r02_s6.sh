#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "videocompare or hash" > gpurun_out/s6_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s6_pytest.log
python scripts/kernel_bench.py --only hsvanim > gpurun_out/s6_kernel_hsv.jsonl 2> gpurun_out/s6_kernel_hsv.err
python - > gpurun_out/s6_hash_timing.jsonl 2> gpurun_out/s6_hash_timing.err <<'PY'
import sys, json, time
sys.path.insert(0, "gst-plugin-rs_b200")
import numpy as np, torch, b200vfx
from b200vfx import synth
ctx = b200vfx.Context(0); ctx.set_stream(torch.cuda.current_stream().cuda_stream)
for (w, h) in ((3840, 2160), (1366, 768), (3841, 2161)):
    f = torch.from_numpy(synth.frame_noise("RGBA", w, h, 1)).cuda()
    for algo in ("mean", "gradient", "vertgradient", "doublegradient", "blockhash"):
        for _ in range(3): ctx.hash_image(algo, "RGBA", w, h, f, 4 * w)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        n = 20
        for _ in range(n): ctx.hash_image(algo, "RGBA", w, h, f, 4 * w)
        torch.cuda.synchronize()
        print(json.dumps({"hash_image": algo, "frame": "%dx%d RGBA device" % (w, h), "us_per_frame_sync_call": round((time.perf_counter() - t0) / n * 1e6, 1)}), flush=True)
PY
ncu --set full --clock-control none --import-source on -k regex:hsv_direct_map -s 2 -c 1 -o gpurun_out/s6_hsvdirect -f python scripts/ncu_target.py --kernel hsvfilter --content ramps --opt hsv_memo=0 --launches 4 > gpurun_out/s6_ncu.log 2>&1
tail -5 gpurun_out/s6_pytest.log; cat gpurun_out/s6_kernel_hsv.jsonl | cut -c1-150; cat gpurun_out/s6_hash_timing.jsonl
