"""colorlut fused with the all-gather of the row tiles (include/b200vfx.h: b200vfx_colorlut_process_tile_gather,
SURVEY 8(e) / BASELINE config 5) against the CPU oracle.  On a one-GPU box the N>1 protocol (entry/exit handshake,
epochs, time-outs) is exercised with several "ranks" sharing the device; the real multi-process / multi-GPU run is
scripts/tile_gather_check.py under torchrun (profiles/r01_tile_gather_n*.txt)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
import b200vfx  # noqa: E402
import oracle_binding as orc  # noqa: E402
from b200vfx import sharding, synth  # noqa: E402


def _cube(size=17, kind="mix"):
    return orc.cube_parse(synth.cube_text_3d(size, kind))


PATH = {"stg": 0, "tma": 1, "stg256": 0}   # stg256: 32-byte stores where the geometry allows, 16-byte otherwise


def _ctx(cube, dev=0, stream=None, path="stg"):
    c = b200vfx.Context(dev)
    c.set_option("tile_gather_path", PATH[path])
    c.set_option("tile_gather_cfg", 1 if path == "stg256" else 0)
    c.colorlut_set_lut(cube.kind, cube.size, cube.values, cube.scale, cube.offset)
    if stream is not None:
        c.set_stream(stream.cuda_stream)
    return c


def _to_np(ptr, nbytes, dev=0):
    buf = sharding.DeviceBuffer(ptr, (nbytes,))
    return torch.as_tensor(buf, device="cuda:%d" % dev).cpu().numpy()


@pytest.mark.parametrize("path", ["stg", "tma", "stg256"])
@pytest.mark.parametrize("w,h,pad", [(256, 64, 0), (1920, 33, 0), (250, 17, 0), (251, 9, 12), (64, 5, 64), (3840, 24, 0), (7680, 8, 0)])
def test_world1_matches_oracle(w, h, pad, path):
    cube = _cube()
    stride = 4 * w + pad
    frame = np.zeros((h, stride), np.uint8)
    frame[:, :4 * w] = synth.frame_noise("RGBA", w, h, 77)
    exp = orc.colorlut_apply(cube, "RGBA", w, h, frame)
    with _ctx(cube, path=path) as ctx:
        pf = sharding.PeerFrames(ctx, None, h, stride, nbuf=2)
        d_in = torch.from_numpy(frame).cuda()
        for _ in range(3):
            k = pf.process(w, d_in, stride)
            assert pf.status() == 0
            got = torch.as_tensor(pf.frame(k), device="cuda").cpu().numpy()
            assert (got[:, :4 * w] == exp[:, :4 * w]).all()
            assert not got[:, 4 * w:].any()          # padding bytes of the frame buffer are never written
        pf.close()


@pytest.mark.parametrize("path", ["stg", "tma", "stg256"])
def test_world1_lut1d(path):
    cube = orc.cube_parse(synth.cube_text_1d(16))
    w, h = 512, 16
    frame = synth.frame_noise("RGBA", w, h, 5)
    exp = orc.colorlut_apply(cube, "RGBA", w, h, frame)
    with _ctx(cube, path=path) as ctx:
        pf = sharding.PeerFrames(ctx, None, h, 4 * w, nbuf=1)
        k = pf.process(w, torch.from_numpy(frame).cuda(), 4 * w)
        assert pf.status() == 0
        assert (torch.as_tensor(pf.frame(k), device="cuda").cpu().numpy() == exp).all()
        pf.close()


def _virtual_ranks(world, w, h, devices, epochs=4, nbuf=1, align=1, path="stg"):
    """`world` ranks in ONE process (rank r on devices[r]): raw peer_alloc pointers are shared directly"""
    cube = _cube()
    stride = 4 * w
    ctxs, streams = [], []
    for r in range(world):
        with torch.cuda.device(devices[r]):
            streams.append(torch.cuda.Stream(device=devices[r]))
        ctxs.append(_ctx(cube, devices[r], streams[r], path))
    for r in range(world):
        for q in range(world):
            ctxs[r].peer_enable_access(devices[q])
    frames = [[ctxs[r].peer_alloc(h * stride)[0] for r in range(world)] for _ in range(nbuf)]
    flags = [ctxs[r].peer_alloc(256)[0] for r in range(world)]
    for e in range(1, epochs + 1):
        frame = synth.frame_noise("RGBA", w, h, 1000 + e)
        exp = orc.colorlut_apply(cube, "RGBA", w, h, frame)
        tiles = []
        for r in range(world):
            r0, r1 = sharding.row_range(h, world, r, align)
            tiles.append(torch.from_numpy(frame[r0:r1].copy()).to("cuda:%d" % devices[r]))
        torch.cuda.synchronize()
        k = e % nbuf
        for r in range(world):          # all launches are asynchronous: the kernels meet on the device(s)
            r0, r1 = sharding.row_range(h, world, r, align)
            ctxs[r].colorlut_process_tile_gather("RGBA", w, r1 - r0, tiles[r], stride, world, r, frames[k], stride, r0, flags, e)
        for r in range(world):
            assert ctxs[r].peer_status(flags[r]) == 0
            got = _to_np(frames[k][r], h * stride, devices[r]).reshape(h, stride)
            assert (got == exp).all(), "epoch %d rank %d" % (e, r)
    for r in range(world):
        for k in range(nbuf):
            ctxs[r].peer_free(frames[k][r])
        ctxs[r].peer_free(flags[r])
        ctxs[r].close()


@pytest.mark.parametrize("path", ["stg", "tma", "stg256"])
@pytest.mark.parametrize("world,w,h", [(2, 256, 8), (3, 100, 10), (4, 64, 3), (8, 128, 16)])
def test_virtual_ranks_one_device(world, w, h, path):
    # small frames: the kernels of all "ranks" are co-resident on the one GPU, so the handshake cannot starve
    _virtual_ranks(world, w, h, [0] * world, epochs=4, nbuf=1, path=path)


@pytest.mark.parametrize("path", ["stg", "tma"])
def test_virtual_ranks_uneven_and_empty_tiles(path):
    _virtual_ranks(4, 96, 5, [0] * 4, epochs=3, nbuf=2, path=path)    # rows 2+2+1+0: the last rank only takes part in the handshake


@pytest.mark.skipif(b200vfx.device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("path", ["stg", "tma"])
def test_two_devices_one_process(path):
    _virtual_ranks(2, 3840, 540, [0, 1], epochs=4, nbuf=1, path=path)
    n = min(b200vfx.device_count(), 8)
    _virtual_ranks(n, 1920, 256, list(range(n)), epochs=3, nbuf=2, path=path)


@pytest.mark.parametrize("path", ["stg", "tma"])
def test_missing_peer_times_out_instead_of_hanging(path):
    cube = _cube()
    w, h = 128, 8
    with _ctx(cube, path=path) as ctx:
        ctx.set_option("peer_timeout_ms", 50)
        frames = [ctx.peer_alloc(h * 4 * w)[0] for _ in range(2)]
        flags = [ctx.peer_alloc(256)[0] for _ in range(2)]
        tile = torch.from_numpy(synth.frame_noise("RGBA", w, h // 2, 3)).cuda()
        ctx.colorlut_process_tile_gather("RGBA", w, h // 2, tile, 4 * w, 2, 0, frames, 4 * w, 0, flags, 7)   # rank 1 never shows up
        assert ctx.peer_status(flags[0]) == 7
        assert not _to_np(frames[1], h * 4 * w).any()    # nothing was pushed to the peer that never said READY
        for p in frames + flags:
            ctx.peer_free(p)


def test_argument_errors():
    cube = _cube()
    with _ctx(cube) as ctx:
        f, g = ctx.peer_alloc(4096)[0], ctx.peer_alloc(256)[0]
        t = torch.zeros(4096, dtype=torch.uint8, device="cuda")
        with pytest.raises(b200vfx.B200VfxError):
            ctx.colorlut_process_tile_gather("RGBA", 16, 4, t, 64, 1, 0, [f], 64, 0, [g], 0)          # epoch 0
        with pytest.raises(b200vfx.B200VfxError):
            ctx.colorlut_process_tile_gather("RGBA64_LE", 16, 4, t, 128, 1, 0, [f], 128, 0, [g], 1)   # 16-bit: unsupported
        with pytest.raises(b200vfx.B200VfxError):
            ctx.colorlut_process_tile_gather("RGBA", 16, 4, t, 64, 1, 1, [f], 64, 0, [g], 1)          # rank out of range
        with pytest.raises(b200vfx.B200VfxError):
            ctx.colorlut_process_tile_gather("RGBA", 16, 4, t, 32, 1, 0, [f], 64, 0, [g], 1)          # stride < row
        ctx.colorlut_set_mode(1)
        with pytest.raises(b200vfx.B200VfxError):
            ctx.colorlut_process_tile_gather("RGBA", 16, 4, t, 64, 1, 0, [f], 64, 0, [g], 1)          # direct mode
        ctx.peer_free(f); ctx.peer_free(g)
    with b200vfx.Context(0) as ctx:
        f, g = ctx.peer_alloc(4096)[0], ctx.peer_alloc(256)[0]
        with pytest.raises(b200vfx.B200VfxError, match="No LUT"):
            ctx.colorlut_process_tile_gather("RGBA", 16, 4, f, 64, 1, 0, [f], 64, 0, [g], 1)
        ctx.peer_free(f); ctx.peer_free(g)
