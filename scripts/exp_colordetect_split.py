#!/usr/bin/env python
"""colordetect, quality 10: all SMs per launch against 1/2, 1/3, 1/4 of them (consecutive launches of a train then overlap);
device time per launch from a CUDA-graph replay"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gst-plugin-rs_b200")); sys.path.insert(0, os.path.join(ROOT, "scripts")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, b200vfx
import oracle_binding as orc
from b200vfx import synth
from kernel_bench import graph_time
W, H = 3840, 2160
ctx = b200vfx.Context(0)
for cname, gen in (("noise", lambda i: synth.frame_noise("RGBA", W, H, 100 + i)), ("ramps", lambda i: np.ascontiguousarray(np.roll(synth.frame_ramps("RGBA", W, H), 4 * 131 * i, axis=1))), ("natural", lambda i: synth.frame_natural("RGBA", W, H, 200 + i, amp=3))):
    host = [gen(i) for i in range(8)]
    fr = [torch.from_numpy(f).cuda() for f in host]
    hist = torch.zeros(32768, dtype=torch.int32, device="cuda")
    for q in (10, 4, 2, 1):
        exp = orc.colordetect_histogram("RGBA", W, H, host[7], q)
        res = {}
        for split in (1, 2, 3, 4):
            ctx.set_option("cd_split", split)
            def fn(i):
                ctx.set_stream(torch.cuda.current_stream().cuda_stream)
                ctx.colordetect_histogram("RGBA", W, H, fr[i % 8], 4 * W, q, hist)
            res["split%d_us" % split] = round(graph_time(fn) * 1e6, 2)
            torch.cuda.synchronize()
            assert (hist.cpu().numpy().view(np.uint32) == exp).all(), (cname, q, split)     # 16 launches: the last one read frame 7
        print(json.dumps({"content": cname, "quality": q, **res}), flush=True)
ctx.close()
