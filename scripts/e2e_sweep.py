#!/usr/bin/env python
"""End-to-end (pinned host -> pinned host) variants of colorlut on a 4K RGBA frame: staged copy-engine pipeline,
full zero-copy TMA kernel, and the two hybrids (one direction by copy engine, the other by the kernel).
Each variant is measured twice, interleaved, to expose run-to-run modes.  One JSON per line."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gst-plugin-rs_b200"))
import numpy as np
import torch

import b200vfx
from b200vfx import synth

W, H = 3840, 2160
ctx = b200vfx.Context(0)
k, s, v, sc, of = b200vfx.cube_parse(synth.cube_text_3d(33, "mix"))
ctx.colorlut_set_lut(k, s, v, sc, of)
srcs = [torch.from_numpy(synth.frame_noise("RGBA", W, H, 100 + i) if i % 2 else np.ascontiguousarray(np.roll(synth.frame_ramps("RGBA", W, H), 4 * 97 * i, axis=1))).pin_memory() for i in range(4)]
dsts = [torch.empty_like(t).pin_memory() for t in srcs]
ctx.set_option("zero_copy", 0)
ctx.colorlut_process("RGBA", W, H, srcs[0].numpy(), 4 * W, dsts[0].numpy(), 4 * W)
ref = dsts[0].clone()


def measure(n=24):
    for i in range(4):
        ctx.colorlut_process("RGBA", W, H, srcs[i % 4].numpy(), 4 * W, dsts[i % 4].numpy(), 4 * W)
    t0 = time.perf_counter()
    for i in range(n):
        ctx.colorlut_process("RGBA", W, H, srcs[i % 4].numpy(), 4 * W, dsts[i % 4].numpy(), 4 * W)
    return (time.perf_counter() - t0) / n


variants = [
    ("staged chunk=auto", {"zero_copy": 0, "zc_hybrid": 0}, 0),
    ("staged chunk=540", {"zero_copy": 0, "zc_hybrid": 0}, 540),
    ("zero-copy TMA cfg2 grid64", {"zero_copy": 1, "zc_hybrid": 0, "zc_cfg": 2, "zc_ctas": 1, "zc_grid": 64}, 0),
    ("zero-copy TMA cfg2 grid148", {"zero_copy": 1, "zc_hybrid": 0, "zc_cfg": 2, "zc_ctas": 1, "zc_grid": 0}, 0),
    ("hybrid: DMA in, kernel stores out, chunk=auto", {"zero_copy": 1, "zc_hybrid": 1}, 0),
    ("hybrid: DMA in, kernel stores out, chunk=540", {"zero_copy": 1, "zc_hybrid": 1}, 540),
    ("hybrid: DMA in, kernel stores out, chunk=2160", {"zero_copy": 1, "zc_hybrid": 1}, 2160),
    ("hybrid: kernel loads in, DMA out, chunk=auto", {"zero_copy": 1, "zc_hybrid": 2}, 0),
    ("hybrid: kernel loads in, DMA out, chunk=540", {"zero_copy": 1, "zc_hybrid": 2}, 540),
    ("auto (probe + watchdog)", {"zero_copy": 2, "zc_hybrid": 0, "zc_cfg": 2, "zc_ctas": 1, "zc_grid": 64}, 0),
]
for rep in range(2):
    for name, opts, rows in variants:
        for o, val in opts.items():
            ctx.set_option(o, val)
        ctx.set_chunk_rows(rows)
        dsts[0].zero_()
        ctx.colorlut_process("RGBA", W, H, srcs[0].numpy(), 4 * W, dsts[0].numpy(), 4 * W)
        same = bool((dsts[0] == ref).all())
        t = measure()
        print(json.dumps({"variant": name, "rep": rep, "ms": round(t * 1e3, 4), "fps": round(1 / t, 1), "pcie_GBps_each_way": round(W * H * 4 / t / 1e9, 2), "identical": same}), flush=True)
ctx.close()
