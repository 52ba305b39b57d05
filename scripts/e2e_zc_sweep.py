#!/usr/bin/env python
"""zero-copy e2e kernel: stream-kernel variant (tile bytes / stages / threads) x grid size, pinned 4K RGBA frames"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gst-plugin-rs_b200"))
import numpy as np, torch, b200vfx
from b200vfx import synth
W, H = 3840, 2160
ctx = b200vfx.Context(0)
k, s, v, sc, of = b200vfx.cube_parse(synth.cube_text_3d(33, "mix")); ctx.colorlut_set_lut(k, s, v, sc, of)
srcs = [torch.from_numpy(synth.frame_noise("RGBA", W, H, 100 + i) if i % 2 else np.ascontiguousarray(np.roll(synth.frame_ramps("RGBA", W, H), 4 * 97 * i, axis=1))).pin_memory() for i in range(4)]
dsts = [torch.empty_like(t).pin_memory() for t in srcs]
def measure(n=16):
    for i in range(3): ctx.colorlut_process("RGBA", W, H, srcs[i % 4].numpy(), 4 * W, dsts[i % 4].numpy(), 4 * W)
    t0 = time.perf_counter()
    for i in range(n): ctx.colorlut_process("RGBA", W, H, srcs[i % 4].numpy(), 4 * W, dsts[i % 4].numpy(), 4 * W)
    return (time.perf_counter() - t0) / n
ctx.set_option("zero_copy", 1)
names = {0: "16K x4 x256", 1: "8K x4 x256", 2: "4K x4 x256", 3: "4K x8 x128", 4: "2K x8 x128", 5: "8K x6 x512", 6: "4K x6 x512", 7: "2K x6 x256"}
for cfg in (2, 1, 0, 3, 4, 6, 7):
    row = {"cfg": cfg, "tile x stages x threads": names[cfg]}
    for grid in (32, 64, 96, 148):
        for ctas in (1, 2):
            if ctas == 2 and grid != 148: continue
            ctx.set_option("zc_cfg", cfg); ctx.set_option("zc_grid", grid * ctas); ctx.set_option("zc_ctas", ctas)
            t = measure()
            row["grid%d" % (grid * ctas)] = round(1 / t, 1)
    print(json.dumps(row), flush=True)
ctx.close()
