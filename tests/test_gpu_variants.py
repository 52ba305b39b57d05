"""GPU tests for kernel variants that must be bit-identical to each other and to the oracle."""
import numpy as np
import pytest

import b200vfx
import oracle_binding as orc
from b200vfx import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = b200vfx.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("tma", [1, 0])
@pytest.mark.parametrize("w,h,pad", [(3840, 2160, 0), (7680, 4320, 0), (1920, 1080, 64), (64, 48, 0), (32, 8, 16), (4096, 16, 0)])
def test_blockhash_tma_and_plain_variants(ctx, tma, w, h, pad):
    torch = pytest.importorskip("torch")
    ctx.set_option("blockhash_tma", tma)
    try:
        frame = synth.frame_noise("RGBA", w, h, 0x5EED0004, stride=4 * w + pad)
        frame[::5, 3:4 * w:28] = 0   # transparent pixels count as 765
        exp = orc.blockhash_sums("RGBA", w, h, frame)
        sums = np.zeros(64, np.uint32)
        ctx.blockhash_sums("RGBA", w, h, frame, frame.shape[1], sums)     # host frame
        assert (sums == exp).all()
        d = torch.from_numpy(frame).cuda()
        ds = torch.zeros(64, dtype=torch.int32, device="cuda")
        ctx.set_stream(torch.cuda.current_stream().cuda_stream)
        for _ in range(3):                                                 # device frame, repeated (sums are re-zeroed)
            ctx.blockhash_sums("RGBA", w, h, d, frame.shape[1], ds)
        torch.cuda.synchronize()
        assert (ds.cpu().numpy().view(np.uint32) == exp).all()
    finally:
        ctx.set_option("blockhash_tma", 1)


def test_contexts_on_every_visible_device():
    """one context per GPU in one process (function attributes / streams / tables are per device)"""
    torch = pytest.importorskip("torch")
    n = torch.cuda.device_count()
    cube1 = orc.cube_parse(synth.cube_text_1d(64, 2.0))     # 1D LUT -> TMA streaming kernel (needs the smem attribute)
    cube3 = orc.cube_parse(synth.cube_text_3d(9, "mix"))
    w, h = 1024, 96
    frame = synth.frame_natural("RGBA", w, h, 4)
    for dev in range(n):
        with b200vfx.Context(dev) as c:
            for cube in (cube1, cube3):
                c.colorlut_set_lut(cube.kind, cube.size, cube.values, cube.scale, cube.offset)
                exp = orc.colorlut_apply(cube, "RGBA", w, h, frame)
                out = np.zeros_like(frame)
                c.colorlut_process("RGBA", w, h, frame, 4 * w, out, 4 * w)              # host path
                assert (out == exp).all(), dev
                d_in = torch.from_numpy(frame).to("cuda:%d" % dev)
                d_out = torch.zeros_like(d_in)
                c.colorlut_process("RGBA", w, h, d_in, 4 * w, d_out, 4 * w)             # device path on the ctx's own stream
                c.synchronize()
                assert (d_out.cpu().numpy() == exp).all(), dev


def test_videocompare_config4_two_4k_streams(ctx):
    """BASELINE config 4: stream 0 = frame A, stream 1 = frame A with 1 % of the pixels perturbed"""
    w, h = 3840, 2160
    a = synth.frame_ramps("RGBA", w, h)
    b = a.copy()
    idx = synth.pcg32(w * h // 100, 0x5EED0004) % np.uint32(w * h)
    b.reshape(-1, 4)[idx, :3] ^= 0x80
    sa, sb = np.zeros(64, np.uint32), np.zeros(64, np.uint32)
    ctx.blockhash_sums("RGBA", w, h, a, 4 * w, sa)
    ctx.blockhash_sums("RGBA", w, h, b, 4 * w, sb)
    assert (sa == orc.blockhash_sums("RGBA", w, h, a)).all() and (sb == orc.blockhash_sums("RGBA", w, h, b)).all()
    ba, bb = b200vfx.blockhash_bits(sa, w, h), b200vfx.blockhash_bits(sb, w, h)
    assert (ba == orc.blockhash_bits(sa, w, h)).all() and (bb == orc.blockhash_bits(sb, w, h)).all()
    assert b200vfx.hash_distance(ba, ba) == 0 and 0 <= b200vfx.hash_distance(ba, bb) <= 8
