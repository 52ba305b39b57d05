import json, os, sys, time
sys.path.insert(0, "gst-plugin-rs_b200")
import numpy as np, torch, b200vfx
from b200vfx import synth
W, H = 3840, 2160
ctx = b200vfx.Context(0)
k, s, v, sc, of = b200vfx.cube_parse(synth.cube_text_3d(33, "mix")); ctx.colorlut_set_lut(k, s, v, sc, of)
srcs = [torch.from_numpy(synth.frame_noise("RGBA", W, H, 100 + i)).pin_memory() for i in range(4)]
dsts = [torch.empty_like(t).pin_memory() for t in srcs]
ctx.set_option("zero_copy", 0)
for rows in (540, 270, 180, 135, 90, 68, 45, 34):
    ctx.set_chunk_rows(rows)
    for i in range(3): ctx.colorlut_process("RGBA", W, H, srcs[i % 4].numpy(), 4 * W, dsts[i % 4].numpy(), 4 * W)
    t0 = time.perf_counter()
    for i in range(16): ctx.colorlut_process("RGBA", W, H, srcs[i % 4].numpy(), 4 * W, dsts[i % 4].numpy(), 4 * W)
    t = (time.perf_counter() - t0) / 16
    print(json.dumps({"staged chunk_rows": rows, "chunks": -(-H // rows), "fps": round(1 / t, 1), "GBps_each_way": round(W * H * 4 / t / 1e9, 2)}), flush=True)
