"""A/B timing of the table-lookup kernels: plain lane-consecutive gathers (memo_tile=0) vs per-tile shared-memory
sub-cube copy (memo_tile=1), 4K RGBA, device-resident, CUDA events."""
import sys, os, json
sys.path.insert(0, "gst-plugin-rs_b200")
import numpy as np, torch, b200vfx
from b200vfx import synth
W, H = 3840, 2160
ctx = b200vfx.Context(0); ctx.set_stream(torch.cuda.current_stream().cuda_stream)
k, s, v, sc, of = b200vfx.cube_parse(synth.cube_text_3d(33, "mix")); ctx.colorlut_set_lut(k, s, v, sc, of)
gens = (("ramps", lambda i: np.ascontiguousarray(np.roll(synth.frame_ramps("RGBA", W, H), 4 * 131 * i, axis=1))),
        ("noise", lambda i: synth.frame_noise("RGBA", W, H, 100 + i)),
        ("natural3", lambda i: synth.frame_natural("RGBA", W, H, 200 + i, amp=3)),
        ("natural8", lambda i: synth.frame_natural("RGBA", W, H, 300 + i, amp=8)),
        ("natural1", lambda i: synth.frame_natural("RGBA", W, H, 400 + i, amp=1)))
def timeit(fn, n=120):
    for i in range(10): fn(i)
    torch.cuda.synchronize(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True); a.record()
    for i in range(n): fn(i)
    b.record(); torch.cuda.synchronize()
    return round(a.elapsed_time(b) * 1e3 / n, 2)
for name, gen in gens:
    fr = [torch.from_numpy(gen(i)).cuda() for i in range(6)]; out = [torch.empty_like(f) for f in fr]
    res = {"content": name}
    for tile in (0, 1):
        ctx.set_option("memo_tile", tile)
        res["colorlut_tile%d_us" % tile] = timeit(lambda i: ctx.colorlut_process("RGBA", W, H, fr[i % 6], 4 * W, out[i % 6], 4 * W))
    ctx.set_option("hsv_memo", 1)
    for tile in (0, 1):
        ctx.set_option("memo_tile", tile)
        res["hsvfilter_tile%d_us" % tile] = timeit(lambda i: ctx.hsvfilter_process("RGBA", W, H, fr[i % 6], 4 * W, hue_shift=90.0))
    ctx.set_option("hsv_memo", -1)
    print(json.dumps(res), flush=True)
