"""Format conversion kernels and the device-resident chain they enable."""
import numpy as np
import pytest

import b200vfx
import np_convert as npc
import oracle_binding as orc
from b200vfx import synth

pytestmark = pytest.mark.gpu
FMTS = list(npc.PACKED)


@pytest.fixture(scope="module")
def ctx():
    c = b200vfx.Context(0)
    yield c
    c.close()


def noise(fmt, w, h, seed, pad=0):
    bpp = npc.PACKED[fmt][0]
    stride = ((w * bpp + 3) // 4) * 4 + pad
    rng = np.random.default_rng(seed)
    return rng.integers(0, 256, (h, stride), dtype=np.uint8), stride


@pytest.mark.parametrize("src_fmt", FMTS)
@pytest.mark.parametrize("dst_fmt", FMTS)
def test_convert_packed_all_pairs(ctx, src_fmt, dst_fmt):
    for (w, h, pad) in ((64, 9, 0), (37, 5, 8), (1281, 3, 4)):
        frame, ss = noise(src_fmt, w, h, 3 * w + h, pad)
        dbpp = npc.PACKED[dst_fmt][0]
        ds = ((w * dbpp + 3) // 4) * 4 + pad
        exp = npc.convert_packed(src_fmt, dst_fmt, w, h, frame, ds, fill=0x5A)
        out = np.full((h, ds), 0x5A, np.uint8)
        ctx.convert_packed(src_fmt, dst_fmt, w, h, frame, ss, out, ds)
        assert (out == exp).all(), (src_fmt, dst_fmt, w, h)            # incl. untouched padding


def test_convert_packed_4k_device(ctx):
    torch = pytest.importorskip("torch")
    w, h = 3840, 2160
    frame = synth.frame_noise("BGRx", w, h, 5)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    d = torch.from_numpy(frame).cuda()
    for dst_fmt in ("RGBA", "ARGB", "RGB"):
        bpp = npc.PACKED[dst_fmt][0]
        o = torch.zeros((h, w * bpp), dtype=torch.uint8, device="cuda")
        for _ in range(3):
            ctx.convert_packed("BGRx", dst_fmt, w, h, d, 4 * w, o, w * bpp)
        torch.cuda.synchronize()
        assert (o.cpu().numpy() == npc.convert_packed("BGRx", dst_fmt, w, h, frame)).all(), dst_fmt


@pytest.mark.parametrize("in_fmt,out_fmt", [("BGRx", "RGBA"), ("ARGB", "BGRA"), ("xRGB", "xBGR"), ("RGBA", "RGBx"), ("ABGR", "ABGR"),
                                            ("BGRA", "ARGB"), ("RGBx", "ABGR")])
@pytest.mark.parametrize("kind", ["3d", "1d"])
def test_colorlut_fused_convert(ctx, in_fmt, out_fmt, kind):
    """colorlut with the surrounding videoconverts folded in == convert -> colorlut (oracle) -> convert"""
    torch = pytest.importorskip("torch")
    cube = orc.cube_parse(synth.cube_text_3d(17, "mix") if kind == "3d" else synth.cube_text_1d(256, 2.2))
    ctx.colorlut_set_lut(cube.kind, cube.size, cube.values, cube.scale, cube.offset)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    for (w, h) in ((640, 33), (1918, 64)):
        frame, ss = noise(in_fmt, w, h, 17 + w)
        rgba = npc.convert_packed(in_fmt, "RGBA", w, h, frame)
        lut = orc.colorlut_apply(cube, "RGBA", w, h, rgba)
        _, _, _, _, has_a = npc.unpack(in_fmt, w, h, frame)
        if not (has_a and npc.PACKED[out_fmt][4] >= 0):
            lut[:, 3::4] = 255
        exp = npc.convert_packed("RGBA", out_fmt, w, h, lut)
        out = np.zeros((h, 4 * w), np.uint8)
        ctx.colorlut_process_fmt(in_fmt, out_fmt, w, h, frame, ss, out, 4 * w)            # host frames
        assert (out == exp).all(), (in_fmt, out_fmt, w, h, "host")
        d, o = torch.from_numpy(frame).cuda(), torch.zeros((h, 4 * w), dtype=torch.uint8, device="cuda")
        for _ in range(2):
            ctx.colorlut_process_fmt(in_fmt, out_fmt, w, h, d, ss, o, 4 * w)              # device frames
        torch.cuda.synchronize()
        assert (o.cpu().numpy() == exp).all(), (in_fmt, out_fmt, w, h, "device")


@pytest.mark.parametrize("fmt", ["RGBA", "BGRx", "RGB", "ARGB"])
@pytest.mark.parametrize("w,h,kind", [(64, 48, 0), (641, 361, 601), (1920, 1080, 0), (33, 17, 709), (1280, 720, 709)])
def test_planar_round_trip_against_spec(ctx, fmt, w, h, kind):
    torch = pytest.importorskip("torch")
    frame, ss = noise(fmt, w, h, w + h)
    with_a = npc.PACKED[fmt][4] >= 0
    exp = npc.to_planar(fmt, w, h, frame, kind, with_alpha=with_a)
    cw, ch = (w + 1) // 2, (h + 1) // 2
    ys, cs = ((w + 3) // 4) * 4, ((cw + 3) // 4) * 4
    planes = [np.full((h, ys), 7, np.uint8), np.full((ch, cs), 7, np.uint8), np.full((ch, cs), 7, np.uint8)]
    strides = [ys, cs, cs]
    if with_a:
        planes.append(np.full((h, ys), 7, np.uint8)); strides.append(ys)
    ctx.convert_to_planar(fmt, "A420" if with_a else "I420", w, h, frame, ss, planes, strides, kind)   # host planes
    for p, e in zip(planes, exp):
        assert (p[:, :e.shape[1]] == e).all() and (p[:, e.shape[1]:] == 7).all(), (fmt, w, h, kind)
    # device planes, and back to packed
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    dsrc = torch.from_numpy(frame).cuda()
    dpl = [torch.zeros(p.shape, dtype=torch.uint8, device="cuda") for p in planes]
    ctx.convert_to_planar(fmt, "A420" if with_a else "I420", w, h, dsrc, ss, dpl, strides, kind)
    bpp = npc.PACKED[fmt][0]
    dback = torch.zeros((h, w * bpp), dtype=torch.uint8, device="cuda")
    ctx.convert_from_planar("A420" if with_a else "I420", fmt, w, h, dpl, strides, dback, w * bpp, kind)
    torch.cuda.synchronize()
    for p, e in zip(dpl, exp):
        assert (p.cpu().numpy()[:, :e.shape[1]] == e).all()
    assert (dback.cpu().numpy() == npc.from_planar(exp, fmt, w, h, kind)).all(), (fmt, w, h, kind)
    back = np.zeros((h, w * bpp), np.uint8)
    ctx.convert_from_planar("A420" if with_a else "I420", fmt, w, h, planes, strides, back, w * bpp, kind)   # host
    assert (back == npc.from_planar(exp, fmt, w, h, kind)).all()
    # sanity of the specification itself: a round trip stays within a few code values
    r0, g0, b0, _, _ = npc.unpack(fmt, w, h, frame)
    gray = np.repeat(np.arange(16, 236, dtype=np.uint8), 3)                       # greys survive almost exactly
    gr = np.tile(gray, (2, 1))
    gp = npc.to_planar("RGB", gr.shape[1] // 3, 2, gr, kind)
    assert np.abs(npc.from_planar(gp, "RGB", gr.shape[1] // 3, 2, kind).astype(int) - gr).max() <= 3


def _planar_frame(w, h, seed, with_a, pad=0, natural=False):
    """random (or smooth) I420 / A420 planes with GStreamer's strides (App. E) + optional extra padding"""
    rng = np.random.default_rng(seed)
    cw, ch = (w + 1) // 2, (h + 1) // 2
    ys, cs = ((w + 3) // 4) * 4 + pad, ((cw + 3) // 4) * 4 + pad
    shapes = [(h, ys), (ch, cs), (ch, cs)] + ([(h, ys)] if with_a else [])
    if natural:
        yy, xx = np.mgrid[0:h, 0:ys]
        base = [((xx * 200 // max(1, ys)) + (yy * 40 // max(1, h)) + 16).astype(np.uint8)]
        base += [np.full((ch, cs), v, np.uint8) for v in (110, 150)]
        planes = [np.clip(b.astype(np.int64) + rng.integers(-3, 4, b.shape), 0, 255).astype(np.uint8) for b in base]
        if with_a:
            planes.append(rng.integers(0, 256, (h, ys), dtype=np.uint8))
    else:
        planes = [rng.integers(0, 256, sh, dtype=np.uint8) for sh in shapes]
    return planes, [sh[1] for sh in shapes]


def _planar_expected(cube, fmt, w, h, planes, kind):
    """the composition the fused kernel replaces: I420 -> RGBA (spec), colorlut (oracle), RGBA -> I420 (spec)"""
    with_a = fmt == "A420"
    rgba = npc.from_planar(planes, "RGBA", w, h, kind)
    lut = orc.colorlut_apply(cube, "RGBA", w, h, rgba)
    return npc.to_planar("RGBA", w, h, lut, kind, with_alpha=with_a)


@pytest.mark.parametrize("lut", ["3d", "1d"])
@pytest.mark.parametrize("fmt,w,h,kind,pad", [("I420", 64, 48, 0, 0), ("I420", 641, 361, 601, 0), ("A420", 1920, 1080, 0, 0), ("I420", 33, 17, 709, 0),
                                              ("A420", 1280, 720, 709, 12), ("I420", 8, 2, 0, 0), ("I420", 1, 1, 0, 0), ("I420", 1288, 6, 0, 4)])
def test_colorlut_on_planar_frames(ctx, lut, fmt, w, h, kind, pad):
    """colorlut on I420 / A420 with both converts fused == convert -> colorlut (oracle) -> convert"""
    torch = pytest.importorskip("torch")
    cube = orc.cube_parse(synth.cube_text_3d(17, "mix") if lut == "3d" else synth.cube_text_1d(256, 2.2))
    ctx.colorlut_set_lut(cube.kind, cube.size, cube.values, cube.scale, cube.offset)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    planes, strides = _planar_frame(w, h, 5 * w + h, fmt == "A420", pad)
    exp = _planar_expected(cube, fmt, w, h, planes, kind)
    out = [np.full(p.shape, 9, np.uint8) for p in planes]
    ctx.colorlut_process_planar(fmt, w, h, planes, strides, out, strides, kind)                      # host planes
    for o, e in zip(out, exp):
        assert (o[:, :e.shape[1]] == e).all() and (o[:, e.shape[1]:] == 9).all(), (fmt, w, h, "host")
    dsrc = [torch.from_numpy(p).cuda() for p in planes]
    ddst = [torch.full(p.shape, 9, dtype=torch.uint8, device="cuda") for p in planes]
    for _ in range(2):
        ctx.colorlut_process_planar(fmt, w, h, dsrc, strides, ddst, strides, kind)                   # device planes
    torch.cuda.synchronize()
    for o, e in zip(ddst, exp):
        o = o.cpu().numpy()
        assert (o[:, :e.shape[1]] == e).all() and (o[:, e.shape[1]:] == 9).all(), (fmt, w, h, "device")
    # and it is what the three separate device calls produce
    rgba = torch.zeros((h, 4 * w), dtype=torch.uint8, device="cuda")
    graded = torch.zeros_like(rgba)
    back = [torch.zeros(p.shape, dtype=torch.uint8, device="cuda") for p in planes]
    ctx.convert_from_planar(fmt, "RGBA", w, h, dsrc, strides, rgba, 4 * w, kind)
    ctx.colorlut_process("RGBA", w, h, rgba, 4 * w, graded, 4 * w)
    ctx.convert_to_planar("RGBA", fmt, w, h, graded, 4 * w, back, strides, kind)
    torch.cuda.synchronize()
    for o, b, e in zip(ddst, back, exp):
        assert (o.cpu().numpy()[:, :e.shape[1]] == b.cpu().numpy()[:, :e.shape[1]]).all()


def test_colorlut_on_planar_4k_and_errors(ctx):
    torch = pytest.importorskip("torch")
    w, h = 3840, 2160
    cube = orc.cube_parse(synth.cube_text_3d(33, "mix"))
    ctx.colorlut_set_lut(cube.kind, cube.size, cube.values, cube.scale, cube.offset)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    for natural in (False, True):
        planes, strides = _planar_frame(w, h, 77, False, natural=natural)
        exp = _planar_expected(cube, "I420", w, h, planes, 0)
        dsrc = [torch.from_numpy(p).cuda() for p in planes]
        ddst = [torch.zeros(p.shape, dtype=torch.uint8, device="cuda") for p in planes]
        ctx.colorlut_process_planar("I420", w, h, dsrc, strides, ddst, strides)
        torch.cuda.synchronize()
        for o, e in zip(ddst, exp):
            assert (o.cpu().numpy() == e).all()
    with pytest.raises(b200vfx.B200VfxError):
        ctx.colorlut_process_planar("I420", w, h, dsrc, strides, dsrc, strides)                       # in place
    with pytest.raises(b200vfx.B200VfxError):
        ctx.colorlut_process_planar("RGBA", w, h, dsrc, strides, ddst, strides)                       # not planar
    with pytest.raises(b200vfx.B200VfxError):
        ctx.colorlut_process_planar("I420", w, h, dsrc, strides, [p.cpu().numpy() for p in ddst], strides)   # device -> host planes
    with pytest.raises(b200vfx.B200VfxError):
        ctx.colorlut_process_planar("I420", w, h, dsrc, [s - 8 for s in strides], ddst, strides)     # stride < row
    with b200vfx.Context(0) as fresh:
        with pytest.raises(b200vfx.B200VfxError) as e:
            fresh.colorlut_process_planar("I420", w, h, dsrc, strides, ddst, strides)
        assert e.value.code == b200vfx.ERR_NOT_NEGOTIATED                                              # imp.rs:210-213


def test_device_resident_chain_with_converters():
    """BGRx camera frame -> [fused convert] colorlut -> hsvdetector -> RGBA -> I420 -> roundedcorners (A420), everything in
    HBM on one stream: one upload, one download per plane.  The RGB part must equal the oracle chain bit for bit; the I420
    part equals the conversion spec."""
    w, h = 1920, 1080
    src = synth.frame_natural("BGRx", w, h, 99, amp=5)
    cube = orc.cube_parse(synth.cube_text_3d(33, "mix"))
    dkw = dict(hue_ref=120.0, hue_var=60.0, saturation_ref=0.6, saturation_var=0.4, value_ref=0.6, value_var=0.4)
    # oracle chain
    rgba = npc.convert_packed("BGRx", "RGBA", w, h, src)
    lut = orc.colorlut_apply(cube, "RGBA", w, h, rgba, threads=8)
    lut[:, 3::4] = 255                                                          # BGRx has no alpha -> RGBx padding 255
    det = orc.hsvdetector("RGBx", "RGBA", w, h, lut, hue_ref=120.0, hue_var=60.0, sat_ref=0.6, sat_var=0.4, val_ref=0.6, val_var=0.4, threads=8)
    yuv = npc.to_planar("RGBA", w, h, det, 0)
    mask = orc.roundmask(w, h, w, 64)
    with b200vfx.Context(0) as ctx:
        ctx.colorlut_set_lut(cube.kind, cube.size, cube.values, cube.scale, cube.offset)
        nb = 4 * w * h
        d_in, d_lut, d_det = ctx.device_alloc(nb), ctx.device_alloc(nb), ctx.device_alloc(nb)
        cw, ch = w // 2, h // 2
        d_y, d_u, d_v, d_a = ctx.device_alloc(w * h), ctx.device_alloc(cw * ch), ctx.device_alloc(cw * ch), ctx.device_alloc(w * ((h + 1) // 2 * 2))
        d_out = [ctx.device_alloc(w * h), ctx.device_alloc(cw * ch), ctx.device_alloc(cw * ch), ctx.device_alloc(w * h)]
        ctx.roundmask_generate(w, h, w, 64, d_a)                                # once per caps / radius
        got = [np.zeros((h, w), np.uint8), np.zeros((ch, cw), np.uint8), np.zeros((ch, cw), np.uint8), np.zeros((h, w), np.uint8)]
        got_det = np.zeros((h, 4 * w), np.uint8)
        for _ in range(3):
            ctx.upload(d_in, 4 * w, src, 4 * w, 4 * w, h)
            ctx.colorlut_process_fmt("BGRx", "RGBx", w, h, d_in, 4 * w, d_lut, 4 * w)
            ctx.hsvdetector_process("RGBx", "RGBA", w, h, d_lut, 4 * w, d_det, 4 * w, **dkw)
            ctx.convert_to_planar("RGBA", "I420", w, h, d_det, 4 * w, [d_y, d_u, d_v], [w, cw, cw], 0)
            ctx.a420_append(w, h, [d_y, d_u, d_v], [w, cw, cw], d_a, w, d_out, [w, cw, cw, w])
            ctx.download(got_det, 4 * w, d_det, 4 * w, 4 * w, h)
            for i, (g, rb, rows) in enumerate(zip(got, (w, cw, cw, w), (h, ch, ch, h))):
                ctx.download(g, rb, d_out[i], rb, rb, rows)
            ctx.synchronize()
            assert (got_det == det).all()
            for g, e in zip(got[:3], yuv):
                assert (g == e).all()
            assert (got[3] == mask[:h]).all()
        for p in [d_in, d_lut, d_det, d_y, d_u, d_v, d_a] + d_out:
            ctx.device_free(p)
