#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "rgba64 or blockhash or videocompare or hash or convert or colorlut_fused" > gpurun_out/s13_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s13_pytest.log
python scripts/kernel_bench.py --only colorlut64,videofx > gpurun_out/s13_kernel.jsonl 2> gpurun_out/s13_kernel.err
python - > gpurun_out/s13_hash_timing.jsonl 2> gpurun_out/s13_hash_timing.err <<'PY'
import sys, json, time
sys.path.insert(0, "gst-plugin-rs_b200")
import numpy as np, torch, b200vfx
from b200vfx import synth
ctx = b200vfx.Context(0); ctx.set_stream(torch.cuda.current_stream().cuda_stream)
for (w, h) in ((3840, 2160), (1366, 768)):
    f = torch.from_numpy(synth.frame_noise("RGBA", w, h, 1)).cuda()
    for algo in ("mean", "gradient", "vertgradient", "doublegradient", "blockhash"):
        for _ in range(3): ctx.hash_image(algo, "RGBA", w, h, f, 4 * w)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        n = 20
        for _ in range(n): ctx.hash_image(algo, "RGBA", w, h, f, 4 * w)
        torch.cuda.synchronize()
        print(json.dumps({"hash_image": algo, "frame": "%dx%d RGBA device" % (w, h), "us_per_frame_sync_call": round((time.perf_counter() - t0) / n * 1e6, 1)}), flush=True)
k, s, v, sc, of = b200vfx.cube_parse(synth.cube_text_3d(33, "mix")); ctx.colorlut_set_lut(k, s, v, sc, of)
W, H = 3840, 2160
fr = [torch.from_numpy(np.ascontiguousarray(np.roll(synth.frame_ramps("BGRx", W, H), 4 * 131 * i, axis=1))).cuda() for i in range(6)]
out = [torch.empty_like(x) for x in fr]
def t(fn, n=100):
    for i in range(10): fn(i)
    torch.cuda.synchronize(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True); a.record()
    for i in range(n): fn(i)
    b.record(); torch.cuda.synchronize(); return round(a.elapsed_time(b) * 1e3 / n, 2)
print(json.dumps({"colorlut_fused_convert_bgrx_to_rgba_us": t(lambda i: ctx.colorlut_process_fmt("BGRx", "RGBA", W, H, fr[i % 6], 4 * W, out[i % 6], 4 * W)),
                  "colorlut_rgba_us": t(lambda i: ctx.colorlut_process("RGBA", W, H, fr[i % 6], 4 * W, out[i % 6], 4 * W)),
                  "convert_packed_bgrx_to_rgba_us": t(lambda i: ctx.convert_packed("BGRx", "RGBA", W, H, fr[i % 6], 4 * W, out[i % 6], 4 * W)), "content": "ramps 4K"}), flush=True)
y = torch.empty((H, W), dtype=torch.uint8, device="cuda"); u = torch.empty((H // 2, W // 2), dtype=torch.uint8, device="cuda"); vv = torch.empty_like(u)
print(json.dumps({"convert_rgba_to_i420_us": t(lambda i: ctx.convert_to_planar("RGBA", "I420", W, H, fr[i % 6], 4 * W, [y, u, vv], [W, W // 2, W // 2], 0)),
                  "convert_i420_to_rgba_us": t(lambda i: ctx.convert_from_planar("I420", "RGBA", W, H, [y, u, vv], [W, W // 2, W // 2], out[i % 6], 4 * W, 0)), "frame": "4K"}), flush=True)
PY
tail -4 gpurun_out/s13_pytest.log; cut -c1-230 gpurun_out/s13_kernel.jsonl | head -12; cat gpurun_out/s13_hash_timing.jsonl; tail -3 gpurun_out/s13_hash_timing.err
