import sys, os, json
sys.path.insert(0, "gst-plugin-rs_b200")
import numpy as np, torch, b200vfx
from b200vfx import synth
W, H, R = 3840, 2160, 12
ctx = b200vfx.Context(0); ctx.set_stream(torch.cuda.current_stream().cuda_stream)
k, s, v, sc, of = b200vfx.cube_parse(synth.cube_text_3d(33, "mix")); ctx.colorlut_set_lut(k, s, v, sc, of)
ramps = lambda i: np.ascontiguousarray(np.roll(synth.frame_ramps("RGBA", W, H), 4 * 97 * i, axis=1))
noise = lambda i: synth.frame_noise("RGBA", W, H, 100 + i)
frames = [ramps(i // 2) if i % 2 == 0 else noise(i // 2) for i in range(R)]
fr = [torch.from_numpy(f).cuda() for f in frames]; out = [torch.empty_like(f) for f in fr]
def timeit(fn, n=240):
    for i in range(24): fn(i)
    torch.cuda.synchronize(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True); a.record()
    for i in range(n): fn(i)
    b.record(); torch.cuda.synchronize()
    return round(a.elapsed_time(b) * 1e3 / n, 2)
res = {"lib": os.path.basename(os.environ.get("B200VFX_LIB", "default"))}
for ctas in (4, 2, 8):
    ctx.set_option("memo_ctas", ctas)
    res["mix_ctas%d_us" % ctas] = timeit(lambda i: ctx.colorlut_process("RGBA", W, H, fr[i % R], 4 * W, out[i % R], 4 * W))
ctx.set_option("memo_ctas", 4)
res["noise_only_us"] = timeit(lambda i: ctx.colorlut_process("RGBA", W, H, fr[(2 * i + 1) % R], 4 * W, out[(2 * i + 1) % R], 4 * W))
print(json.dumps(res), flush=True)
