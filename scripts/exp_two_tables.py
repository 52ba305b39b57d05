"""Two memoised elements in one pipeline: colorlut (64 MiB answer table) -> hsvfilter (64 MiB answer table), 4K RGBA frames
that stay in HBM.  128 MiB of tables exceed the 126 MB L2: what does the chain cost against the two elements alone?"""
import sys, json
sys.path.insert(0, "gst-plugin-rs_b200")
import numpy as np, torch, b200vfx
from b200vfx import synth
W, H = 3840, 2160
st = torch.cuda.current_stream().cuda_stream
a, b = b200vfx.Context(0), b200vfx.Context(0)
a.set_stream(st); b.set_stream(st)
k, s, v, sc, of = b200vfx.cube_parse(synth.cube_text_3d(33, "mix")); a.colorlut_set_lut(k, s, v, sc, of)
b.set_option("hsv_memo", 1)
def timeit(fn, n=120):
    for i in range(10): fn(i)
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True); e0.record()
    for i in range(n): fn(i)
    e1.record(); torch.cuda.synchronize()
    return round(e0.elapsed_time(e1) * 1e3 / n, 2)
gens = (("ramps", lambda i: np.ascontiguousarray(np.roll(synth.frame_ramps("RGBA", W, H), 4 * 131 * i, axis=1))),
        ("natural3", lambda i: synth.frame_natural("RGBA", W, H, 200 + i, amp=3)),
        ("noise", lambda i: synth.frame_noise("RGBA", W, H, 100 + i)))
for name, gen in gens:
    fr = [torch.from_numpy(gen(i)).cuda() for i in range(6)]; out = [torch.empty_like(f) for f in fr]
    lut = timeit(lambda i: a.colorlut_process("RGBA", W, H, fr[i % 6], 4 * W, out[i % 6], 4 * W))
    hsv = timeit(lambda i: b.hsvfilter_process("RGBA", W, H, out[i % 6], 4 * W, hue_shift=90.0))
    def chain(i):
        a.colorlut_process("RGBA", W, H, fr[i % 6], 4 * W, out[i % 6], 4 * W)
        b.hsvfilter_process("RGBA", W, H, out[i % 6], 4 * W, hue_shift=90.0)
    ch = timeit(chain)
    print(json.dumps({"content": name, "colorlut_alone_us": lut, "hsvfilter_alone_us": hsv, "sum_us": round(lut + hsv, 2), "chain_us_per_frame": ch}), flush=True)
