"""run-to-run spread of the A/B stream: fresh context each repetition, memo_ctas given on the command line"""
import sys, os, json
sys.path.insert(0, "gst-plugin-rs_b200")
import numpy as np, torch, b200vfx
from b200vfx import synth
W, H, R = 3840, 2160, 12
ctas = int(sys.argv[1]); reps = int(sys.argv[2])
ramps = lambda i: np.ascontiguousarray(np.roll(synth.frame_ramps("RGBA", W, H), 4 * 97 * i, axis=1))
noise = lambda i: synth.frame_noise("RGBA", W, H, 100 + i)
frames = [ramps(i // 2) if i % 2 == 0 else noise(i // 2) for i in range(R)]
fr = [torch.from_numpy(f).cuda() for f in frames]; out = [torch.empty_like(f) for f in fr]
k, s, v, sc, of = b200vfx.cube_parse(synth.cube_text_3d(33, "mix"))
vals = []
for rep in range(reps):
    ctx = b200vfx.Context(0); ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    ctx.colorlut_set_lut(k, s, v, sc, of); ctx.set_option("memo_ctas", ctas)
    for i in range(64 * 5): ctx.colorlut_process("RGBA", W, H, fr[i % R], 4 * W, out[i % R], 4 * W)
    torch.cuda.synchronize(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True); a.record()
    n = 64 * 30
    for i in range(n): ctx.colorlut_process("RGBA", W, H, fr[i % R], 4 * W, out[i % R], 4 * W)
    b.record(); torch.cuda.synchronize()
    vals.append(round(a.elapsed_time(b) * 1e3 / n, 2)); ctx.close()
print(json.dumps({"memo_ctas": ctas, "us_per_frame": vals}))
