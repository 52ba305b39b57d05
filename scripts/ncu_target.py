#!/usr/bin/env python
"""Tiny launcher for `ncu --set full`: runs a handful of launches of ONE kernel on ONE content."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gst-plugin-rs_b200"))
import numpy as np
import torch

import b200vfx
from b200vfx import synth

ap = argparse.ArgumentParser()
ap.add_argument("--kernel", default="memo", choices=["memo", "direct", "direct64", "hsvfilter", "hsvdetector", "blockhash", "colordetect", "hash", "fmt", "planar"])
ap.add_argument("--content", default="ramps", choices=["ramps", "noise", "natural"])
ap.add_argument("--lut", type=int, default=33)
ap.add_argument("--launches", type=int, default=6)
ap.add_argument("--quality", type=int, default=1)
ap.add_argument("--memo-tile", type=int, default=0)
ap.add_argument("--opt", action="append", default=[], help="name=value context option, repeatable (e.g. hsv_memo=0)")
a = ap.parse_args()
W, H = 3840, 2160
ctx = b200vfx.Context(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
ctx.set_option("memo_tile", a.memo_tile)
for kv in a.opt:
    name, val = kv.split("=")
    ctx.set_option(name, int(val))


def frame(fmt, w, h, i):
    if a.content == "ramps":
        return np.ascontiguousarray(np.roll(synth.frame_ramps(fmt, w, h), 8 * 131 * i, axis=1))
    if a.content == "natural":
        return synth.frame_natural(fmt, w, h, 200 + i)
    return synth.frame_noise(fmt, w, h, 100 + i)


if a.kernel in ("memo", "direct", "direct64"):
    k, s, v, sc, of = b200vfx.cube_parse(synth.cube_text_3d(a.lut, "mix"))
    ctx.colorlut_set_lut(k, s, v, sc, of)
    ctx.colorlut_set_mode(0 if a.kernel == "memo" else 1)
    fmt = "RGBA64_LE" if a.kernel == "direct64" else "RGBA"
    bpp = 8 if a.kernel == "direct64" else 4
    fr = [torch.from_numpy(frame(fmt, W, H, i)).cuda() for i in range(4)]
    out = [torch.empty_like(f) for f in fr]
    for i in range(a.launches):
        ctx.colorlut_process(fmt, W, H, fr[i % 4], bpp * W, out[i % 4], bpp * W)
elif a.kernel == "hsvfilter":
    fr = [torch.from_numpy(frame("RGBA", W, H, i)).cuda() for i in range(4)]
    for i in range(a.launches):
        ctx.hsvfilter_process("RGBA", W, H, fr[i % 4], 4 * W, hue_shift=90.0)
elif a.kernel == "hsvdetector":
    w, h = 1920, 1080
    fr = [torch.from_numpy(frame("BGRx", w, h, i)).cuda() for i in range(4)]
    out = [torch.empty_like(f) for f in fr]
    for i in range(a.launches):
        ctx.hsvdetector_process("BGRx", "RGBA", w, h, fr[i % 4], 4 * w, out[i % 4], 4 * w, hue_ref=120.0, hue_var=30.0,
                                saturation_ref=0.8, saturation_var=0.2, value_ref=0.8, value_var=0.2)
elif a.kernel == "hash":        # videocompare mean hash: grayscale + Lanczos3 resize kernels
    fr = [torch.from_numpy(frame("RGBA", W, H, i)).cuda() for i in range(2)]
    for i in range(a.launches):
        ctx.hash_image("mean", "RGBA", W, H, fr[i % 2], 4 * W)
elif a.kernel == "fmt":         # colorlut with the surrounding converts fused in (BGRx -> RGBA)
    k, s, v, sc, of = b200vfx.cube_parse(synth.cube_text_3d(a.lut, "mix"))
    ctx.colorlut_set_lut(k, s, v, sc, of)
    fr = [torch.from_numpy(frame("BGRx", W, H, i)).cuda() for i in range(4)]
    out = [torch.empty_like(f) for f in fr]
    for i in range(a.launches):
        ctx.colorlut_process_fmt("BGRx", "RGBA", W, H, fr[i % 4], 4 * W, out[i % 4], 4 * W)
elif a.kernel == "planar":      # colorlut on I420 frames, both converts fused in
    k, s, v, sc, of = b200vfx.cube_parse(synth.cube_text_3d(a.lut, "mix"))
    ctx.colorlut_set_lut(k, s, v, sc, of)
    rgba = [frame("RGBA", W, H, i) for i in range(4)]
    strides = [W, W // 2, W // 2]
    src = []
    for f in rgba:   # planes of a frame with the chosen content: luma = green channel, chroma = subsampled red / blue
        px = f.reshape(H, W, 4)
        src.append([torch.from_numpy(np.ascontiguousarray(px[:, :, 1])).cuda(), torch.from_numpy(np.ascontiguousarray(px[::2, ::2, 0])).cuda(),
                    torch.from_numpy(np.ascontiguousarray(px[::2, ::2, 2])).cuda()])
    dst = [[torch.empty_like(p) for p in fr_] for fr_ in src]
    for i in range(a.launches):
        ctx.colorlut_process_planar("I420", W, H, src[i % 4], strides, dst[i % 4], strides)
elif a.kernel == "colordetect":
    fr = [torch.from_numpy(frame("RGBA", W, H, i)).cuda() for i in range(4)]
    hist = torch.zeros(32768, dtype=torch.int32, device="cuda")
    for i in range(a.launches):
        ctx.colordetect_histogram("RGBA", W, H, fr[i % 4], 4 * W, a.quality, hist)
else:
    fr = [torch.from_numpy(frame("RGBA", W, H, i)).cuda() for i in range(4)]
    sums = torch.zeros(64, dtype=torch.int32, device="cuda")
    for i in range(a.launches):
        ctx.blockhash_sums("RGBA", W, H, fr[i % 4], 4 * W, sums)
torch.cuda.synchronize()
ctx.close()
