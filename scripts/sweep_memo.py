#!/usr/bin/env python
"""Sweep the colorlut RGBA memo-kernel variants (plain LDG kernel PX, TMA stream cfg/CTAs/hint) on the three
contents, next to a same-size device copy (the practical ceiling for a 66 MB kernel).  One JSON per line."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gst-plugin-rs_b200"))
import numpy as np, torch
import b200vfx
from b200vfx import synth

W, H, RING = 3840, 2160, 6
PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def timeit(fn, iters=50, warm=5):
    for i in range(warm): fn(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(iters): fn(i)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e-3 / iters


contents = {
    "ramps": lambda i: np.ascontiguousarray(np.roll(synth.frame_ramps("RGBA", W, H), 4 * 131 * i, axis=1)),
    "natural": lambda i: synth.frame_natural("RGBA", W, H, 200 + i, amp=3),
    "noise": lambda i: synth.frame_noise("RGBA", W, H, 100 + i),
}
ctx = b200vfx.Context(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
k, s, v, sc, of = b200vfx.cube_parse(synth.cube_text_3d(33, "mix"))
ctx.colorlut_set_lut(k, s, v, sc, of)
data = {}
for name, fn in contents.items():
    fr = [torch.from_numpy(fn(i)).cuda() for i in range(RING)]
    data[name] = (fr, [torch.empty_like(f) for f in fr])
fr, out = data["noise"]
t = timeit(lambda i: out[i % RING].copy_(fr[i % RING]))
print(json.dumps({"variant": "torch_copy_33MB", "us": round(t * 1e6, 2), "frac": round(2 * W * H * 4 / t / 1e9 / PEAK, 4)}), flush=True)
# (the L2-persisting access window is deliberately not swept here: enabling it once raises a device-wide set-aside that
#  slows every later variant -- see profiles/r01_l2_persist_experiment.jsonl)
variants = [("plain", {"stream_path": 0, "memo_px": px, "pdl": pdl}) for px in (4, 8) for pdl in (0, 1)]
for cfg in (0, 1, 2, 7):
    for pdl in (0, 1):
        variants.append(("tma", {"stream_path": 1, "stream_cfg": cfg, "stream_ctas": 0, "stream_hint": 1, "pdl": pdl}))
for kind, opts in variants:
    for o, val in opts.items(): ctx.set_option(o, val)
    row = {"variant": kind, **opts}
    for name, (fr, out) in data.items():
        t = timeit(lambda i: ctx.colorlut_process("RGBA", W, H, fr[i % RING], 4 * W, out[i % RING], 4 * W))
        row[name + "_us"] = round(t * 1e6, 2)
    row["ramps_frac"] = round(2 * W * H * 4 / (row["ramps_us"] * 1e-6) / 1e9 / PEAK, 4)
    print(json.dumps(row), flush=True)
ctx.close()
