#!/usr/bin/env python
"""End-to-end colorlut on pinned host frames (4K RGBA, 33^3): synchronous calls against the asynchronous host-frame mode,
zero-copy kernel against the copy-engine pipeline, chunk sizes, frames in flight; plus the plain concurrent-memcpy ceiling."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gst-plugin-rs_b200"))
import numpy as np, torch, b200vfx
from b200vfx import synth
W, H = 3840, 2160
k, s, v, sc, of = b200vfx.cube_parse(synth.cube_text_3d(33, "mix"))
frames = [synth.frame_ramps("RGBA", W, H), synth.frame_noise("RGBA", W, H, 1), synth.frame_natural("RGBA", W, H, 2), synth.frame_noise("RGBA", W, H, 3)]
RING = 8
h_in = [torch.from_numpy(frames[i % 4]).pin_memory() for i in range(RING)]
h_out = [torch.empty_like(t).pin_memory() for t in h_in]
np_in = [t.numpy() for t in h_in]; np_out = [t.numpy() for t in h_out]

def measure(zero_copy, chunk_rows, inflight, n=96):
    ctx = b200vfx.Context(0)
    ctx.colorlut_set_lut(k, s, v, sc, of)
    ctx.set_option("zero_copy", zero_copy)
    ctx.set_chunk_rows(chunk_rows)
    for i in range(8):   # warm-up (table build, probe calls)
        ctx.colorlut_process("RGBA", W, H, np_in[i % RING], 4 * W, np_out[i % RING], 4 * W)
    if inflight > 0:
        ctx.set_host_async(True)
    fences = []
    t0 = time.perf_counter()
    for i in range(n):
        if inflight > 0 and len(fences) >= inflight:
            f = fences.pop(0); f.wait(); f.close()
        ctx.colorlut_process("RGBA", W, H, np_in[i % RING], 4 * W, np_out[i % RING], 4 * W)
        if inflight > 0:
            fences.append(ctx.fence())
    ctx.synchronize()
    dt = (time.perf_counter() - t0) / n
    for f in fences:
        f.close()
    ctx.close()
    return round(1.0 / dt, 1)

sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
dd_in = torch.empty((H, 4 * W), dtype=torch.uint8, device="cuda"); dd_out = torch.empty_like(dd_in)
def both():
    with torch.cuda.stream(sa): dd_in.copy_(h_in[0], non_blocking=True)
    with torch.cuda.stream(sb): h_out[1].copy_(dd_out, non_blocking=True)
for _ in range(3): both()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(30): both()
torch.cuda.synchronize()
print(json.dumps({"concurrent_pinned_memcpy_ceiling_fps": round(30 / (time.perf_counter() - t0), 1)}), flush=True)
for zc, rows, infl in ((2, 0, 0), (0, 0, 0), (1, 0, 0), (1, 0, 3), (2, 0, 1), (2, 0, 2), (2, 0, 3), (2, 0, 4), (2, 0, 5), (0, 1080, 3), (0, 540, 3), (2, 0, 3)):
    print(json.dumps({"zero_copy": zc, "chunk_rows": rows or "auto", "frames_in_flight": infl or "synchronous calls",
                      "frames_per_s": measure(zc, rows, infl)}), flush=True)
