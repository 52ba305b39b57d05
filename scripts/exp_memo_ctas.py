"""persistent table-lookup kernel: CTAs per SM (option memo_ctas) vs throughput of a stream of 4K frames (device-resident,
ring of 12 in + 12 out buffers, CUDA events): with fewer CTAs per SM the next frame's kernel (PDL) runs beside the current one"""
import sys, os, json
sys.path.insert(0, "gst-plugin-rs_b200")
import numpy as np, torch, b200vfx
from b200vfx import synth
W, H, R = 3840, 2160, 12
ctx = b200vfx.Context(0); ctx.set_stream(torch.cuda.current_stream().cuda_stream)
k, s, v, sc, of = b200vfx.cube_parse(synth.cube_text_3d(33, "mix")); ctx.colorlut_set_lut(k, s, v, sc, of)
ramps = lambda i: np.ascontiguousarray(np.roll(synth.frame_ramps("RGBA", W, H), 4 * 97 * i, axis=1))
noise = lambda i: synth.frame_noise("RGBA", W, H, 100 + i)
nat = lambda i: synth.frame_natural("RGBA", W, H, 200 + i, amp=3)
sets = {"mix A/B": [ramps(i // 2) if i % 2 == 0 else noise(i // 2) for i in range(R)], "ramps": [ramps(i) for i in range(R)],
        "noise": [noise(i) for i in range(R)]}
print(json.dumps({"lib": os.environ.get("B200VFX_LIB", "default")}))
def timeit(fn, n=240):
    for i in range(24): fn(i)
    torch.cuda.synchronize(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True); a.record()
    for i in range(n): fn(i)
    b.record(); torch.cuda.synchronize()
    return round(a.elapsed_time(b) * 1e3 / n, 2)
for name, frames in sets.items():
    fr = [torch.from_numpy(f).cuda() for f in frames]; out = [torch.empty_like(f) for f in fr]
    res = {"content": name}
    for ctas, px in ((8, 8), (4, 8), (2, 8), (4, 16), (2, 16), (4, 4)):
        ctx.set_option("memo_ctas", ctas); ctx.set_option("memo_px", px)
        res["ctas%d_px%d_us" % (ctas, px)] = timeit(lambda i: ctx.colorlut_process("RGBA", W, H, fr[i % R], 4 * W, out[i % R], 4 * W))
    ctx.set_option("pdl", 0); ctx.set_option("memo_ctas", 8); ctx.set_option("memo_px", 8)
    res["ctas8_nopdl_us"] = timeit(lambda i: ctx.colorlut_process("RGBA", W, H, fr[i % R], 4 * W, out[i % R], 4 * W))
    ctx.set_option("pdl", 1)
    print(json.dumps(res), flush=True)
    del fr, out
