#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/s8_pytest.log; cat gpurun_out/s8_pytest.log
python - > gpurun_out/s8_blockhash.txt 2>&1 <<'PY'
import sys, json
sys.path.insert(0, "gst-plugin-rs_b200")
import numpy as np, torch, b200vfx
from b200vfx import synth
ctx = b200vfx.Context(0); ctx.set_stream(torch.cuda.current_stream().cuda_stream)
for (W, H) in ((3840, 2160), (7680, 4320), (1920, 1080)):
    fr = [torch.from_numpy(synth.frame_noise("RGBA", W, H, 5 + i)).cuda() for i in range(4)]
    sums = torch.zeros(64, dtype=torch.int32, device="cuda")
    for tma in (0, 1):
        ctx.set_option("blockhash_tma", tma)
        for i in range(5): ctx.blockhash_sums("RGBA", W, H, fr[i % 4], 4 * W, sums)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(60): ctx.blockhash_sums("RGBA", W, H, fr[i % 4], 4 * W, sums)
        b.record(); torch.cuda.synchronize()
        t = a.elapsed_time(b) / 60 * 1e-3
        print(json.dumps({"kernel": "blockhash_sums", "tma": tma, "frame": "%dx%d" % (W, H), "us": round(t * 1e6, 2), "GBps": round(W * H * 4 / t / 1e9, 1)}))
PY
cat gpurun_out/s8_blockhash.txt
ncu --set full --clock-control none --import-source on -k regex:blockhash -s 3 -c 1 -f -o gpurun_out/s8_blockhash_tma \
    python scripts/ncu_target.py --kernel blockhash --content noise > gpurun_out/s8_ncu.log 2>&1
