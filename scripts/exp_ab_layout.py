#!/usr/bin/env python
"""Why does the A/B stream (bench `value`) vary between 37k and 45k frames/s across boxes while frame A alone and frame B
alone do not?  Repeats the A/B stream inside one process with the allocations shifted (a dummy allocation of varying size
in front of the table / the frames), with fresh contexts, and with the options that change how consecutive frames overlap."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gst-plugin-rs_b200"))
import numpy as np, torch, b200vfx
from b200vfx import synth
W, H, R = 3840, 2160, 12
ramps = lambda i: np.ascontiguousarray(np.roll(synth.frame_ramps("RGBA", W, H), 4 * 97 * i, axis=1))
noise = lambda i: synth.frame_noise("RGBA", W, H, 100 + i)
host = [ramps(i // 2) if i % 2 == 0 else noise(i // 2) for i in range(R)]
k, s, v, sc, of = b200vfx.cube_parse(synth.cube_text_3d(33, "mix"))

def run(ctx, fr, out, n=64 * 20, order=None):
    idx = order or list(range(R))
    for i in range(64 * 4): ctx.colorlut_process("RGBA", W, H, fr[idx[i % R]], 4 * W, out[idx[i % R]], 4 * W)
    torch.cuda.synchronize(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True); a.record()
    for i in range(n): ctx.colorlut_process("RGBA", W, H, fr[idx[i % R]], 4 * W, out[idx[i % R]], 4 * W)
    b.record(); torch.cuda.synchronize()
    return round(a.elapsed_time(b) * 1e3 / n, 2)

for trial in range(5):
    pad_tab = torch.empty(1 + trial * 19 * (1 << 20), dtype=torch.uint8, device="cuda")      # shifts the table
    ctx = b200vfx.Context(0); ctx.set_stream(torch.cuda.current_stream().cuda_stream); ctx.colorlut_set_lut(k, s, v, sc, of)
    d0 = torch.from_numpy(host[0]).cuda(); o0 = torch.empty_like(d0)
    ctx.colorlut_process("RGBA", W, H, d0, 4 * W, o0, 4 * W); torch.cuda.synchronize()        # table allocated + built now
    pad_fr = torch.empty(1 + trial * 7 * (1 << 20), dtype=torch.uint8, device="cuda")         # shifts the frames
    fr = [torch.from_numpy(f).cuda() for f in host]; out = [torch.empty_like(f) for f in fr]
    res = {"trial": trial, "ab_us": [run(ctx, fr, out) for _ in range(3)]}
    res["a_only_us"] = run(ctx, fr, out, order=[0, 2, 4, 6, 8, 10] * 2)
    res["b_only_us"] = run(ctx, fr, out, order=[1, 3, 5, 7, 9, 11] * 2)
    res["aabb_us"] = run(ctx, fr, out, order=[0, 2, 1, 3, 4, 6, 5, 7, 8, 10, 9, 11])
    for ctas in (2, 3, 8):
        ctx.set_option("memo_ctas", ctas); res["ab_ctas%d_us" % ctas] = run(ctx, fr, out)
    ctx.set_option("memo_ctas", 4)
    ctx.set_option("pdl", 0); res["ab_no_pdl_us"] = run(ctx, fr, out); ctx.set_option("pdl", 1)
    print(json.dumps(res), flush=True)
    ctx.close(); del fr, out, pad_tab, pad_fr, d0, o0
    torch.cuda.empty_cache()
