#!/usr/bin/env python
"""Short driver for compute-sanitizer (memcheck / racecheck / synccheck) over the kernels added in round 2: direct hsv map
kernels, single-launch blockhash / colordetect reductions, videocompare resize + fractional blockhash, format converters,
fused colorlut+convert, the small-frame PDL ring.  Small frames, few launches: the tools slow kernels down 10-100x."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gst-plugin-rs_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import b200vfx
from b200vfx import synth

ctx = b200vfx.Context(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
w, h = 640, 360
k, s, v, sc, of = b200vfx.cube_parse(synth.cube_text_3d(9, "mix"))
ctx.colorlut_set_lut(k, s, v, sc, of)
f = synth.frame_noise("RGBA", w, h, 1)
d = torch.from_numpy(f).cuda()
outs = [torch.zeros_like(d) for _ in range(3)]
for i in range(12):                                  # PDL ring of 3 small frames
    ctx.colorlut_process("RGBA", w, h, d, 4 * w, outs[i % 3], 4 * w)
ctx.colorlut_process_fmt("BGRx", "ARGB", w, h, d, 4 * w, outs[0], 4 * w)
f16 = torch.from_numpy(synth.frame_noise("RGBA64_LE", w, h, 2)).cuda()
o16 = torch.zeros_like(f16)
ctx.colorlut_process("RGBA64_LE", w, h, f16, 8 * w, o16, 8 * w)
ctx.set_option("hsv_memo", 0)
x = d.clone()
for i in range(3):
    ctx.hsvfilter_process("RGBA", w, h, x, 4 * w, hue_shift=10.0 * i + 1)
    ctx.hsvdetector_process("BGRx", "RGBA", w, h, d, 4 * w, outs[1], 4 * w, hue_ref=30.0 * i)
sums = torch.zeros(128, dtype=torch.int32, device="cuda")
hist = torch.zeros(32768, dtype=torch.int32, device="cuda")
for i in range(3):
    ctx.blockhash_sums("RGBA", w, h, d, 4 * w, sums)
    ctx.blockhash_sums_batch("RGBA", w, h, [d, outs[0]], [4 * w, 4 * w], sums)
    ctx.colordetect_histogram("RGBA", w, h, d, 4 * w, 1, hist)
    ctx.colordetect_histogram("RGBA", w, h, d, 4 * w, 10, hist)
sums16 = torch.zeros(512, dtype=torch.int32, device="cuda")
for rows in (0, 1, 2):                               # block-sum decompositions, PDL-overlapped trains, 16x8 grid, two frames
    ctx.set_option("blockhash_rows", rows)
    for i in range(4):
        ctx.blockhash_sums_batch("RGBA", w, h, [d, outs[i % 3]], [4 * w, 4 * w], sums16, hw=16, hh=8)
        ctx.blockhash_sums("RGBA", w, h, outs[i % 3], 4 * w, sums)
        ctx.colorlut_process("RGBA", w, h, d, 4 * w, outs[(i + 1) % 3], 4 * w)   # producer of the next frame to hash
ctx.set_option("blockhash_rows", 2)
from b200vfx import sharding
for cfg in (0, 1):                                   # fused tile gather on one rank: 16- and 32-byte stores
    ctx.set_option("tile_gather_cfg", cfg)
    pf = sharding.PeerFrames(ctx, None, h, 4 * w, nbuf=2)
    for i in range(3):
        pf.process(w, d, 4 * w)
    assert pf.status() == 0
    pf.close()
ctx.set_option("tile_gather_cfg", 0)
for algo in ("mean", "gradient", "vertgradient", "doublegradient", "blockhash"):
    ctx.hash_image(algo, "RGBA", w, h, d, 4 * w)
    ctx.hash_image(algo, "RGBA", w - 3, h - 1, d, 4 * w)
big = torch.from_numpy(synth.frame_noise("RGBA", 2051, 1201, 3) | 0x80).cuda()
ctx.hash_image("blockhash", "RGBA", 2051, 1201, big, 4 * 2051)       # sequential f32 chain kernel
ctx.convert_packed("RGBA", "BGRx", w, h, d, 4 * w, outs[2], 4 * w)
rgb = torch.zeros((h, 3 * w), dtype=torch.uint8, device="cuda")
ctx.convert_packed("RGBA", "RGB", w, h, d, 4 * w, rgb, 3 * w)
pl = [torch.zeros((h, w), dtype=torch.uint8, device="cuda"), torch.zeros((h // 2, w // 2), dtype=torch.uint8, device="cuda"),
      torch.zeros((h // 2, w // 2), dtype=torch.uint8, device="cuda"), torch.zeros((h, w), dtype=torch.uint8, device="cuda")]
ctx.convert_to_planar("RGBA", "A420", w, h, d, 4 * w, pl, [w, w // 2, w // 2, w], 0)
ctx.convert_from_planar("A420", "BGRA", w, h, pl, [w, w // 2, w // 2, w], outs[2], 4 * w, 0)
ctx.set_option("memo_tile", 1)
ctx.colorlut_process("RGBA", w, h, d, 4 * w, outs[0], 4 * w)
torch.cuda.synchronize()
ctx.close()
print("racecheck target done")
