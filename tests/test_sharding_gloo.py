"""world_size-2 (and 3) gloo tests of the N>1 path on CPU: row-tile ownership, seams, reassembly by all-gather,
block-sum all-reduce.  The per-tile pixel work is done by the oracle here (there is no GPU in this container);
the GPU variant of the same check is tests/test_gpu_parity.py::test_row_tiles_match_full_frame."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle_binding as orc
from b200vfx import sharding, synth


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, w, h, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cube = orc.cube_parse(synth.cube_text_3d(9, "mix"))
        frame = synth.frame_noise("RGBA", w, h, 0x5EED0005)
        # --- per-pixel path: each rank processes only its row tile -------------------------------------
        r0, r1 = sharding.row_range(h, world, rank)
        tile_in = frame[r0:r1]
        tile_out = orc.colorlut_apply(cube, "RGBA", w, r1 - r0, tile_in) if r1 > r0 else np.zeros((0, 4 * w), np.uint8)
        full = sharding.all_gather_rows(dist, torch.from_numpy(tile_out), h, world).numpy()
        exp = orc.colorlut_apply(cube, "RGBA", w, h, frame)
        ok1 = bool((full == exp).all())
        # seams: rows k*rows_per_gpu +- 1 come from different ranks and must be bit-identical to the 1-GPU result
        per = sharding.row_range(h, world, 0)[1]
        seams = [r for k in range(1, world) for r in (k * per - 1, k * per) if 0 <= r < h]
        ok2 = all((full[r] == exp[r]).all() for r in seams)
        # --- blockhash: tiles aligned to hash-block rows, partial sums all-reduced ----------------------
        bh = h // 8
        b0, b1 = sharding.row_range(h, world, rank, align=bh)
        part = np.zeros(64, np.uint32)
        if b1 > b0:
            nblk = (b1 - b0) // bh
            sub = orc.blockhash_sums("RGBA", w, b1 - b0, frame[b0:b1], hw=8, hh=nblk)
            part[(b0 // bh) * 8:(b0 // bh) * 8 + nblk * 8] = sub
        tot = sharding.all_reduce_sums(dist, torch.from_numpy(part.astype(np.int64))).numpy().astype(np.uint32)
        ok3 = bool((tot == orc.blockhash_sums("RGBA", w, h, frame)).all())
        q.put((rank, ok1, ok2, ok3))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,w,h", [(2, 64, 48), (2, 40, 24), (3, 32, 40)])
def test_row_tile_sharding_gloo(world, w, h):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, w, h, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(r[0] for r in res) == list(range(world))
    assert all(r[1] and r[2] and r[3] for r in res), res


def test_row_range_partitions_every_row_once():
    for h in (1, 7, 48, 1080, 2160, 4320):
        for n in (1, 2, 3, 4, 8):
            for al in (1, 2, max(1, h // 8)):
                seen = np.zeros(h, int)
                for r in range(n):
                    a, b = sharding.row_range(h, n, r, al)
                    assert 0 <= a <= b <= h
                    seen[a:b] += 1
                assert (seen == 1).all()
    assert sharding.row_range(4320, 8, 3) == (1620, 2160)   # 540 rows per GPU (BASELINE config 5)
