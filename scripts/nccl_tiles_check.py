#!/usr/bin/env python
"""N-rank NCCL check of the row-tile path on real GPUs: every rank runs colorlut (device pointers, through the C ABI)
on its row tile of an 8K frame (BASELINE config 5 shape), the tiles are reassembled with ONE all-gather over NVLink,
and every rank compares the whole frame bit for bit with the CPU oracle.  Also all-reduces blockhash partial sums."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gst-plugin-rs_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, torch.distributed as dist
import b200vfx, oracle_binding as orc
from b200vfx import sharding, synth

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
W, H = 7680, 4320
frame = synth.frame_natural("RGBA", W, H, 0x5EED0005)
cube = orc.cube_parse(synth.cube_text_3d(65, "mix"))
ctx = b200vfx.Context(local)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
ctx.colorlut_set_lut(cube.kind, cube.size, cube.values, cube.scale, cube.offset)
r0, r1 = sharding.row_range(H, world, rank)
d_in = torch.from_numpy(frame[r0:r1]).cuda()
d_out = torch.empty_like(d_in)
ctx.colorlut_process("RGBA", W, r1 - r0, d_in, 4 * W, d_out, 4 * W)
full = sharding.all_gather_rows(dist, d_out, H, world)
torch.cuda.synchronize()
exp = orc.colorlut_apply(cube, "RGBA", W, H, frame, threads=16)
ok = bool((full.cpu().numpy() == exp).all())
# blockhash partial sums: tiles aligned to hash-block rows
bh = H // 8
b0, b1 = sharding.row_range(H, world, rank, align=bh)
part = torch.zeros(64, dtype=torch.int32, device="cuda")
if b1 > b0:
    nblk = (b1 - b0) // bh
    sub = torch.zeros(8 * nblk, dtype=torch.int32, device="cuda")
    d_t = torch.from_numpy(frame[b0:b1]).cuda()
    ctx.blockhash_sums("RGBA", W, b1 - b0, d_t, 4 * W, sub, hw=8, hh=nblk)
    part[(b0 // bh) * 8:(b0 // bh) * 8 + 8 * nblk] = sub
tot = sharding.all_reduce_sums(dist, part).cpu().numpy().astype(np.uint32)
ok2 = bool((tot == orc.blockhash_sums("RGBA", W, H, frame)).all())
print("rank %d/%d rows [%d,%d): all-gathered 8K colorlut frame == oracle: %s ; all-reduced blockhash sums == oracle: %s" % (rank, world, r0, r1, ok, ok2), flush=True)
ctx.close()
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok and ok2 else 1)
