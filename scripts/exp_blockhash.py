#!/usr/bin/env python
"""videocompare block sums: one CTA per hash block (round 1) against whole-row streaming CTAs, device time per launch from a
CUDA-graph replay, frames rotating through a ring larger than L2."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gst-plugin-rs_b200")); sys.path.insert(0, os.path.join(ROOT, "scripts"))
import numpy as np, torch
import b200vfx
from b200vfx import synth
from kernel_bench import graph_time, PEAK

ctx = b200vfx.Context(0)
for (W, H) in ((3840, 2160), (1920, 1080), (7680, 4320)):
    ring = 8 if W < 7000 else 4
    frames = [torch.from_numpy(synth.frame_noise("RGBA", W, H, 100 + i)).cuda() for i in range(ring)]
    for hs in (8, 16):
        if W % hs or H % hs:
            continue
        for nfr in (1, 2):
            sums = torch.zeros(hs * hs * nfr, dtype=torch.int32, device="cuda")
            res = {}
            for rows in (0, 1, 2, 3):
                ctx.set_option("blockhash_rows", rows)
                def fn(i):
                    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
                    if nfr == 1:
                        ctx.blockhash_sums("RGBA", W, H, frames[i % ring], 4 * W, sums, hw=hs, hh=hs)
                    else:
                        ctx.blockhash_sums_batch("RGBA", W, H, [frames[(2 * i) % ring], frames[(2 * i + 1) % ring]], [4 * W] * 2, sums, hw=hs, hh=hs)
                t = graph_time(fn)
                res["rows%d" % rows if rows else "blocks"] = round(t * 1e6, 2)
                res["sum_%d" % rows] = int(sums.sum().item())
            assert res["sum_0"] == res["sum_1"] == res["sum_2"] == res["sum_3"]
            gb = nfr * W * H * 4 / (res["rows1"] * 1e-6) / 1e9
            print(json.dumps({"frame": "%dx%d" % (W, H), "hash_grid": hs, "frames_per_launch": nfr, "per_block_ctas_us": res["blocks"],
                              "row_streaming_us": res["rows1"], "row_streaming_1cta_per_sm_us": res["rows2"], "row_streaming_2_per_3sm_us": res["rows3"],
                              "row_streaming_GBps": round(gb, 1), "frac_of_measured_peak": round(gb / PEAK, 3)}), flush=True)
    del frames
ctx.close()
