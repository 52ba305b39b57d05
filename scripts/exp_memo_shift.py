"""timing-only experiment: how fast is the memo kernel on random pixels if the table were 32 / 16 MiB (L2 resident)?
run with B200VFX_LIB=<variant .so>; results of the shifted variants are wrong by construction"""
import sys, os, json
sys.path.insert(0, "gst-plugin-rs_b200")
import numpy as np, torch, b200vfx
from b200vfx import synth
W, H = 3840, 2160
ctx = b200vfx.Context(0); ctx.set_stream(torch.cuda.current_stream().cuda_stream)
k, s, v, sc, of = b200vfx.cube_parse(synth.cube_text_3d(33, "mix")); ctx.colorlut_set_lut(k, s, v, sc, of)
for name, gen in (("ramps", lambda i: np.ascontiguousarray(np.roll(synth.frame_ramps("RGBA", W, H), 4 * 131 * i, axis=1))),
                  ("noise", lambda i: synth.frame_noise("RGBA", W, H, 100 + i)), ("natural", lambda i: synth.frame_natural("RGBA", W, H, 200 + i, amp=3))):
    fr = [torch.from_numpy(gen(i)).cuda() for i in range(6)]; out = [torch.empty_like(f) for f in fr]
    for i in range(10): ctx.colorlut_process("RGBA", W, H, fr[i % 6], 4 * W, out[i % 6], 4 * W)
    torch.cuda.synchronize(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True); a.record()
    for i in range(120): ctx.colorlut_process("RGBA", W, H, fr[i % 6], 4 * W, out[i % 6], 4 * W)
    b.record(); torch.cuda.synchronize()
    print(json.dumps({"lib": os.environ.get("B200VFX_LIB", "default"), "content": name, "us": round(a.elapsed_time(b) * 1e3 / 120, 2)}))
