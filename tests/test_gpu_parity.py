"""GPU parity tests (run on the B200 box): the CUDA path, called through the C ABI, against the CPU
oracle on the same seeded inputs.  Bit-exact everywhere (all outputs are u8/u16/u32)."""
import os

import numpy as np
import pytest

import b200vfx
import oracle_binding as orc
from b200vfx import synth

pytestmark = pytest.mark.gpu
NT = max(1, min(32, os.cpu_count() or 1))


@pytest.fixture(scope="module")
def ctx():
    c = b200vfx.Context(0)
    yield c
    c.close()


def all_colors_frame(fmt="RGBA"):
    idx = np.arange(1 << 24, dtype=np.uint32)
    r, g, b = (idx & 0xFF).astype(np.uint8), ((idx >> 8) & 0xFF).astype(np.uint8), ((idx >> 16) & 0xFF).astype(np.uint8)
    a = (idx * 7 + 3).astype(np.uint8)
    planes = {"R": r, "G": g, "B": b, "A": a}
    order = synth._ORDER[fmt]
    return np.stack([planes[ch] for ch in order], axis=-1).reshape(4096, 4096 * len(order))


def set_cube(ctx, cube):
    ctx.colorlut_set_lut(cube.kind, cube.size, cube.values, cube.scale, cube.offset)


def gpu_colorlut(ctx, fmt, w, h, src, dstride=None, fill=0x5A):
    dstride = dstride or src.shape[1]
    dst = np.full((h, dstride), fill, np.uint8)
    ctx.colorlut_process(fmt, w, h, src, src.shape[1], dst, dstride)
    return dst


# ---- colorlut ------------------------------------------------------------------------------------
@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("n,kind,dom", [(33, "mix", None), (65, "mix", None), (2, "invert", None),
                                        (17, "mix", ((-0.25, 0.0, 0.1), (1.5, 0.75, 0.9)))])
def test_colorlut_rgba_all_colors(ctx, mode, n, kind, dom):
    cube = orc.cube_parse(synth.cube_text_3d(n, kind, domain=dom))
    set_cube(ctx, cube)
    ctx.colorlut_set_mode(mode)
    frame = all_colors_frame("RGBA")
    exp = orc.colorlut_apply(cube, "RGBA", 4096, 4096, frame, threads=NT)
    got = gpu_colorlut(ctx, "RGBA", 4096, 4096, frame)
    ctx.colorlut_set_mode(0)
    assert (got == exp).all()


@pytest.mark.parametrize("mode", [0, 1])
def test_colorlut_1d_rgba_all_values(ctx, mode):
    cube = orc.cube_parse(synth.cube_text_1d(1024, 2.2, domain=((0.0, -0.5, 0.1), (1.0, 1.5, 0.6))))
    set_cube(ctx, cube)
    ctx.colorlut_set_mode(mode)
    frame = all_colors_frame("RGBA")[:512]
    exp = orc.colorlut_apply(cube, "RGBA", 4096, 512, frame, threads=NT)
    got = gpu_colorlut(ctx, "RGBA", 4096, 512, frame)
    ctx.colorlut_set_mode(0)
    assert (got == exp).all()


def test_colorlut_appendix_c_vectors(ctx):
    px = np.array([[95, 130, 194, 1, 217, 207, 235, 2, 15, 163, 33, 3]], np.uint8)
    for n, exp in ((33, [(35, 182, 140), (185, 230, 220), (1, 204, 70)]), (2, [(95, 130, 140), (217, 207, 220), (15, 163, 70)])):
        set_cube(ctx, orc.cube_from_values(3, n, synth.lut_values_3d(n, "mix")))
        for mode in (0, 1):
            ctx.colorlut_set_mode(mode)
            out = gpu_colorlut(ctx, "RGBA", 3, 1, px)[0].reshape(3, 4)
            assert [tuple(p[:3]) for p in out.tolist()] == exp and out[:, 3].tolist() == [1, 2, 3]
    ctx.colorlut_set_mode(0)


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("w,h,spad,dpad,off", [(1, 1, 0, 0, 0), (3, 2, 4, 8, 0), (129, 7, 12, 4, 0), (640, 9, 0, 16, 0),
                                               (257, 5, 3, 5, 1), (1000, 3, 0, 0, 2), (4097, 2, 4, 0, 0)])
def test_colorlut_rgba_strides_and_alignment(ctx, mode, w, h, spad, dpad, off):
    cube = orc.cube_parse(synth.cube_text_3d(9, "mix"))
    set_cube(ctx, cube)
    ctx.colorlut_set_mode(mode)
    sstride, dstride = 4 * w + spad, 4 * w + dpad
    raw = np.full(h * sstride + 8, 0xA5, np.uint8)
    src = raw[off:off + h * sstride].reshape(h, sstride)
    src[:, :4 * w] = synth.frame_noise("RGBA", w, h, 99)[:, :4 * w]
    exp = orc.colorlut_apply(cube, "RGBA", w, h, np.ascontiguousarray(src), dst_stride=dstride)
    rawd = np.full(h * dstride + 8, 0x5A, np.uint8)
    dst = rawd[off:off + h * dstride].reshape(h, dstride)
    ctx.colorlut_process("RGBA", w, h, src.ctypes.data, sstride, dst.ctypes.data, dstride)
    ctx.colorlut_set_mode(0)
    assert (dst == exp).all()          # includes untouched row padding (0x5A)
    assert (rawd[:off] == 0x5A).all() and (rawd[off + h * dstride:] == 0x5A).all()


@pytest.mark.parametrize("fmt", ["RGBA64_LE", "RGBA64_BE"])
@pytest.mark.parametrize("lut", ["3d33", "3d5dom", "1d"])
def test_colorlut_rgba64(ctx, fmt, lut):
    if lut == "3d33":
        cube = orc.cube_parse(synth.cube_text_3d(33, "mix"))
    elif lut == "3d5dom":
        cube = orc.cube_parse(synth.cube_text_3d(5, "mix", domain=((0.0, 0.1, 0.0), (1.0, 0.9, 2.0))))
    else:
        cube = orc.cube_parse(synth.cube_text_1d(4096, 2.2, domain=((0.0, 0.0, 0.0), (0.5, 1.0, 2.0))))
    set_cube(ctx, cube)
    w, h = 1031, 64
    for frame, sstride, dstride in ((synth.frame_noise(fmt, w, h, 77, stride=8 * w + 16), 8 * w + 16, 8 * w + 8),
                                    (synth.frame_ramps(fmt, w, h), 8 * w, 8 * w),
                                    (synth.frame_noise(fmt, w, h, 78, stride=8 * w + 2), 8 * w + 2, 8 * w + 6)):
        exp = orc.colorlut_apply(cube, fmt, w, h, frame, dst_stride=dstride, threads=NT)
        got = gpu_colorlut(ctx, fmt, w, h, frame, dstride)
        assert (got == exp).all()
    # boundary values incl. 0/1/max in every channel
    vals = np.array([0, 1, 2, 255, 256, 32767, 32768, 65534, 65535], np.uint16)
    grid = np.stack(np.meshgrid(vals, vals, vals, vals[:3], indexing="ij"), -1).reshape(1, -1, 4)
    dt = "<u2" if fmt.endswith("LE") else ">u2"
    frame = grid.astype(dt).view(np.uint8).reshape(1, -1)
    wpx = frame.shape[1] // 8
    exp = orc.colorlut_apply(cube, fmt, wpx, 1, frame)
    assert (gpu_colorlut(ctx, fmt, wpx, 1, frame) == exp).all()


def test_colorlut_nan_inf_entries(ctx):
    v = synth.lut_values_3d(5, "mix")
    v[7] = [np.nan, np.inf, -np.inf]
    v[60] = [1e30, -1e30, np.nan]
    cube = orc.cube_from_values(3, 5, v)
    set_cube(ctx, cube)
    frame = all_colors_frame("RGBA")[:1024]
    exp = orc.colorlut_apply(cube, "RGBA", 4096, 1024, frame, threads=NT)
    for mode in (0, 1):
        ctx.colorlut_set_mode(mode)
        assert (gpu_colorlut(ctx, "RGBA", 4096, 1024, frame) == exp).all()
    ctx.colorlut_set_mode(0)
    f16 = synth.frame_noise("RGBA64_LE", 512, 64, 5)
    assert (gpu_colorlut(ctx, "RGBA64_LE", 512, 64, f16) == orc.colorlut_apply(cube, "RGBA64_LE", 512, 64, f16)).all()


def test_colorlut_errors_and_lifecycle(ctx, tmp_path):
    ctx.colorlut_clear()
    src = np.zeros((2, 8), np.uint8)
    with pytest.raises(b200vfx.B200VfxError) as e:
        ctx.colorlut_process("RGBA", 2, 2, src, 8, src.copy(), 8)
    assert e.value.code == b200vfx.ERR_NOT_NEGOTIATED and "No LUT configured" in e.value.msg
    with pytest.raises(b200vfx.B200VfxError) as e:
        ctx.colorlut_load_file(str(tmp_path / "nope.cube"))
    assert e.value.code == b200vfx.ERR_IO
    bad = tmp_path / "bad.cube"
    bad.write_text("LUT_3D_SIZE 2\n0 0 0\n")
    with pytest.raises(b200vfx.B200VfxError) as e:
        ctx.colorlut_load_file(str(bad))
    assert e.value.code == b200vfx.ERR_PARSE and "Failed to parse LUT file" in e.value.msg
    good = tmp_path / "good.cube"
    good.write_text(synth.cube_text_3d(4, "invert"))
    ctx.colorlut_load_file(str(good))
    with pytest.raises(b200vfx.B200VfxError) as e:
        ctx.colorlut_process("BGRA", 2, 2, src, 8, src.copy(), 8)
    assert e.value.code == b200vfx.ERR_UNSUPPORTED
    with pytest.raises(b200vfx.B200VfxError):
        ctx.colorlut_process("RGBA", 4, 2, src, 8, src.copy(), 8)   # stride < row bytes
    ctx.colorlut_process("RGBA", 0, 0, None, 0, None, 0)            # empty frame is a no-op
    px = np.array([[10, 20, 30, 40, 250, 128, 0, 7]], np.uint8)
    out = gpu_colorlut(ctx, "RGBA", 2, 1, px)
    assert out[0].tolist() == [245, 235, 225, 40, 5, 127, 255, 7]


def test_colorlut_device_pointers_and_full_size(ctx):
    torch = pytest.importorskip("torch")
    cube = orc.cube_parse(synth.cube_text_3d(33, "mix"))
    set_cube(ctx, cube)
    w, h = 3840, 2160
    for frame in (synth.frame_ramps("RGBA", w, h), synth.frame_noise("RGBA", w, h, 0x5EED0002)):
        exp = orc.colorlut_apply(cube, "RGBA", w, h, frame, threads=NT)
        d_src = torch.from_numpy(frame).cuda()
        d_dst = torch.zeros_like(d_src)
        ctx.set_stream(torch.cuda.current_stream().cuda_stream)
        for mode in (0, 1):
            ctx.colorlut_set_mode(mode)
            d_dst.zero_()
            ctx.colorlut_process("RGBA", w, h, d_src, 4 * w, d_dst, 4 * w)
            torch.cuda.synchronize()
            assert (d_dst.cpu().numpy() == exp).all()
        ctx.colorlut_set_mode(0)
        # host path (chunked H2D/kernel/D2H pipeline) with a pinned source and pageable destination
        pin = torch.from_numpy(frame).pin_memory()
        assert (gpu_colorlut(ctx, "RGBA", w, h, pin.numpy()) == exp).all()
        # properties at full size: alpha untouched, idempotent under the identity LUT
        assert (exp.reshape(-1, 4)[:, 3] == frame.reshape(-1, 4)[:, 3]).all()
    ident = orc.cube_from_values(3, 33, synth.lut_values_3d(33, "identity"))
    set_cube(ctx, ident)
    frame = synth.frame_noise("RGBA", w, h, 11)
    assert (gpu_colorlut(ctx, "RGBA", w, h, frame) == frame).all()


# ---- hsvfilter -----------------------------------------------------------------------------------
def gpu_hsvfilter(ctx, fmt, w, h, frame, **kw):
    out = frame.copy()
    ctx.hsvfilter_process(fmt, w, h, out, out.shape[1], **kw)
    return out


HSVF = [dict(), dict(hue_shift=90.0), dict(hue_shift=-270.25, saturation_mul=1.7, saturation_off=-0.2, value_mul=0.8, value_off=0.15),
        dict(hue_shift=1e9), dict(saturation_mul=float("nan")), dict(value_off=float("inf"), hue_shift=float("-inf")),
        dict(hue_shift=359.9), dict(hue_shift=-1e-30)]


def _orc_kw(kw):
    m = {"saturation_mul": "sat_mul", "saturation_off": "sat_off", "value_mul": "val_mul", "value_off": "val_off"}
    return {m.get(k, k): v for k, v in kw.items()}


@pytest.mark.parametrize("memo", [0, 1])
@pytest.mark.parametrize("i", range(len(HSVF)))
def test_hsvfilter_all_colors(ctx, i, memo):
    kw = HSVF[i]
    fmt = ["RGBA", "xBGR", "BGRx", "ARGB"][i % 4]
    frame = all_colors_frame(fmt)
    exp = orc.hsvfilter(fmt, 4096, 4096, frame, threads=NT, **_orc_kw(kw))
    ctx.set_option("hsv_memo", memo)
    try:
        assert (gpu_hsvfilter(ctx, fmt, 4096, 4096, frame, **kw) == exp).all()
    finally:
        ctx.set_option("hsv_memo", -1)


def test_hsvfilter_default_is_not_identity(ctx):
    frame = all_colors_frame("RGBA")
    out = gpu_hsvfilter(ctx, "RGBA", 4096, 4096, frame)
    d = out.reshape(-1, 4).astype(np.int16) - frame.reshape(-1, 4).astype(np.int16)
    assert int((d[:, :3] != 0).any(axis=1).sum()) == 11093274 and (d[:, 3] == 0).all()


@pytest.mark.parametrize("memo", [0, 1])
@pytest.mark.parametrize("fmt", ["RGBx", "xRGB", "BGRx", "xBGR", "RGBA", "ARGB", "BGRA", "ABGR", "RGB", "BGR"])
def test_hsvfilter_formats_strides(ctx, fmt, memo):
    kw = dict(hue_shift=123.5, saturation_mul=0.9, saturation_off=0.05, value_mul=1.1, value_off=-0.02)
    ctx.set_option("hsv_memo", memo)
    for (w, h, pad, off) in ((1, 1, 0, 0), (37, 5, 8, 0), (640, 48, 0, 0), (255, 3, 3, 1)):
        bpp = 3 if fmt in ("RGB", "BGR") else 4
        stride = synth.default_stride(fmt, w) + pad
        raw = np.full(h * stride + 8, 0xA5, np.uint8)
        fr = raw[off:off + h * stride].reshape(h, stride)
        fr[:, :bpp * w] = synth.frame_noise(fmt, w, h, 7)[:, :bpp * w]
        exp = orc.hsvfilter(fmt, w, h, np.ascontiguousarray(fr), **_orc_kw(kw))
        ctx.hsvfilter_process(fmt, w, h, fr.ctypes.data, stride, **kw)
        assert (fr == exp).all()
        assert (raw[:off] == 0xA5).all() and (raw[off + h * stride:] == 0xA5).all()
    ctx.set_option("hsv_memo", -1)


def test_hsvfilter_config1_and_device(ctx):
    torch = pytest.importorskip("torch")
    w, h = 640, 480
    for frame in (synth.frame_ramps("RGBA", w, h), synth.frame_noise("RGBA", w, h, 0x5EED0001)):
        exp = orc.hsvfilter("RGBA", w, h, frame, hue_shift=90.0)
        assert (gpu_hsvfilter(ctx, "RGBA", w, h, frame, hue_shift=90.0) == exp).all()
        d = torch.from_numpy(frame).cuda()
        ctx.set_stream(torch.cuda.current_stream().cuda_stream)
        ctx.hsvfilter_process("RGBA", w, h, d, 4 * w, hue_shift=90.0)
        torch.cuda.synchronize()
        assert (d.cpu().numpy() == exp).all()


# ---- hsvdetector ---------------------------------------------------------------------------------
def gpu_hsvdetector(ctx, ifmt, ofmt, w, h, src, dstride=None, **kw):
    dstride = dstride or 4 * w
    dst = np.full((h, dstride), 0x5A, np.uint8)
    ctx.hsvdetector_process(ifmt, ofmt, w, h, src, src.shape[1], dst, dstride, **kw)
    return dst


def _orc_dkw(kw):
    m = {"saturation_ref": "sat_ref", "saturation_var": "sat_var", "value_ref": "val_ref", "value_var": "val_var"}
    return {m.get(k, k): v for k, v in kw.items()}


HSVD = [dict(), dict(hue_ref=120.0, hue_var=30.0, saturation_ref=0.8, saturation_var=0.2, value_ref=0.8, value_var=0.2),
        dict(hue_ref=-700.0, hue_var=180.0, saturation_ref=0.5, saturation_var=0.5, value_ref=0.5, value_var=0.5),
        dict(hue_ref=359.99, hue_var=0.5, saturation_var=1.0, value_var=1.0), dict(hue_ref=float("nan"))]


@pytest.mark.parametrize("memo", [0, 1])
@pytest.mark.parametrize("i", range(len(HSVD)))
def test_hsvdetector_all_colors(ctx, i, memo):
    ctx.set_option("hsv_memo", memo)
    kw = HSVD[i]
    ifmt, ofmt = [("RGBx", "RGBA"), ("BGRx", "ARGB"), ("xRGB", "BGRA"), ("xBGR", "ABGR"), ("RGBx", "ABGR")][i]
    frame = all_colors_frame(ifmt.replace("x", "A"))
    exp = orc.hsvdetector(ifmt, ofmt, 4096, 4096, frame, threads=NT, **_orc_dkw(kw))
    got = gpu_hsvdetector(ctx, ifmt, ofmt, 4096, 4096, frame, **kw)
    assert (got == exp).all()
    if i == 0:
        assert int((got.reshape(-1, 4)[:, 3] == 255).sum()) == 719
    if i == 1:
        assert int((got.reshape(-1, 4)[:, 0] == 255).sum()) == 1415062
    ctx.set_option("hsv_memo", -1)


@pytest.mark.parametrize("memo", [0, 1])
@pytest.mark.parametrize("ifmt", ["RGBx", "xRGB", "BGRx", "xBGR", "RGB", "BGR"])
@pytest.mark.parametrize("ofmt", ["RGBA", "ARGB", "BGRA", "ABGR"])
def test_hsvdetector_format_matrix(ctx, ifmt, ofmt, memo):
    kw = dict(hue_ref=200.0, hue_var=90.0, saturation_ref=0.5, saturation_var=0.5, value_ref=0.5, value_var=0.5)
    ctx.set_option("hsv_memo", memo)
    for (w, h, spad, dpad) in ((1, 1, 0, 0), (37, 5, 8, 12), (1920, 8, 0, 0), (333, 3, 4, 4)):
        src = synth.frame_noise(ifmt, w, h, 0x5EED0003, stride=synth.default_stride(ifmt, w) + spad)
        exp = orc.hsvdetector(ifmt, ofmt, w, h, src, dst_stride=4 * w + dpad, **_orc_dkw(kw))
        assert (gpu_hsvdetector(ctx, ifmt, ofmt, w, h, src, 4 * w + dpad, **kw) == exp).all()
    ctx.set_option("hsv_memo", -1)


@pytest.mark.parametrize("fmt", ["RGB", "BGR"])
def test_hsv_rgb24_memo_device_frames(ctx, fmt):
    """3-byte pixels through the answer tables (map_rgb24_kernel): device-resident frames, packed (flattened to one
    row when 3*w*h is a multiple of 4) and padded strides, widths that are not multiples of 4, in place for hsvfilter
    and 3->4 bytes for hsvdetector; padding bytes must stay untouched."""
    torch = pytest.importorskip("torch")
    fkw = dict(hue_shift=77.0, saturation_mul=1.2, value_off=0.05)
    dkw = dict(hue_ref=200.0, hue_var=90.0, saturation_ref=0.5, saturation_var=0.5, value_ref=0.5, value_var=0.5)
    ctx.set_option("hsv_memo", 1)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    try:
        for (w, h, pad) in ((640, 48, 0), (1000, 3, 0), (333, 4, 1), (4, 1, 0), (2049, 5, 3), (3840, 16, 0), (6, 2, 2), (333, 4, 0), (5, 3, 0)):
            stride = 3 * w + pad
            if stride % 4:
                stride += 4 - stride % 4 if pad else 0
            fr = np.full((h, stride), 0xA5, np.uint8)
            fr[:, :3 * w] = synth.frame_noise(fmt, w, h, 21 + w)[:, :3 * w]
            exp = orc.hsvfilter(fmt, w, h, fr.copy(), **_orc_kw(fkw))
            d = torch.from_numpy(fr).cuda()
            ctx.hsvfilter_process(fmt, w, h, d, stride, **fkw)
            torch.cuda.synchronize()
            assert (d.cpu().numpy() == exp).all(), (w, h, pad)
            for ofmt, dpad in (("RGBA", 0), ("ABGR", 4), ("BGRA", 16)):
                dstride = 4 * w + dpad
                expd = orc.hsvdetector(fmt, ofmt, w, h, fr, dst_stride=dstride, **_orc_dkw(dkw))
                dd = torch.full((h, dstride), 0x5A, dtype=torch.uint8, device="cuda")
                ctx.hsvdetector_process(fmt, ofmt, w, h, torch.from_numpy(fr).cuda(), stride, dd, dstride, **dkw)
                torch.cuda.synchronize()
                assert (dd.cpu().numpy() == expd).all(), (w, h, pad, ofmt)
    finally:
        ctx.set_option("hsv_memo", -1)


def test_hsvdetector_rejects_formats_outside_caps(ctx):
    src = np.zeros((1, 8), np.uint8)
    for ifmt, ofmt in (("RGBA", "RGBA"), ("RGBx", "RGBx"), ("RGBA64_LE", "RGBA"), ("RGB", "BGR")):
        with pytest.raises(b200vfx.B200VfxError) as e:
            ctx.hsvdetector_process(ifmt, ofmt, 1, 1, src, 8, src.copy(), 8)
        assert e.value.code == b200vfx.ERR_UNSUPPORTED


def test_hsvdetector_config3_full_size(ctx):
    w, h = 1920, 1080
    kw = HSVD[1]
    for frame in (synth.frame_ramps("BGRx", w, h), synth.frame_noise("BGRx", w, h, 0x5EED0003)):
        exp = orc.hsvdetector("BGRx", "RGBA", w, h, frame, threads=NT, **_orc_dkw(kw))
        assert (gpu_hsvdetector(ctx, "BGRx", "RGBA", w, h, frame, **kw) == exp).all()


# ---- videocompare / blockhash ----------------------------------------------------------------------
@pytest.mark.parametrize("fmt,w,h,pad", [("RGBA", 64, 48, 0), ("RGBA", 3840, 2160, 0), ("RGBA", 72, 40, 12), ("RGBA", 24, 16, 4),
                                         ("RGB", 64, 48, 0), ("RGB", 1920, 1080, 0), ("RGB", 40, 24, 4), ("RGBA", 8, 8, 0)])
def test_blockhash_sums(ctx, fmt, w, h, pad):
    frame = synth.frame_noise(fmt, w, h, 0x5EED0004, stride=synth.default_stride(fmt, w) + pad)
    if fmt == "RGBA":
        frame[::3, 3:4 * w:16] = 0  # some fully transparent pixels -> counted as 765
    exp = orc.blockhash_sums(fmt, w, h, frame)
    sums = np.full(64, 0xDEADBEEF, np.uint32)
    ctx.blockhash_sums(fmt, w, h, frame, frame.shape[1], sums)
    assert (sums == exp).all()


def test_blockhash_device_and_videocompare_semantics(ctx):
    torch = pytest.importorskip("torch")
    w, h = 3840, 2160
    a = synth.frame_ramps("RGBA", w, h)
    pert = a.copy()
    idx = synth.pcg32(w * h // 100, 0x5EED0004) % np.uint32(w * h)
    pert.reshape(-1, 4)[idx, :3] ^= 0xFF
    red = synth.frame_solid("RGBA", w, h)
    snow = synth.frame_noise("RGBA", w, h, 1)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)

    def bits(frame):
        d = torch.from_numpy(frame).cuda()
        s = torch.zeros(64, dtype=torch.int32, device="cuda")
        ctx.blockhash_sums("RGBA", w, h, d, 4 * w, s)
        torch.cuda.synchronize()
        sums = s.cpu().numpy().view(np.uint32)
        assert (sums == orc.blockhash_sums("RGBA", w, h, frame)).all()
        return b200vfx.blockhash_bits(sums, w, h)

    ba, bp, br, bs = bits(a), bits(pert), bits(red), bits(snow)
    assert b200vfx.hash_distance(ba, ba) == 0
    assert b200vfx.hash_distance(br, bits(red.copy())) == 0       # red vs red -> distance 0 (tests/videocompare.rs:57-103)
    assert b200vfx.hash_distance(br, bs) > 0                      # snow vs red -> no match (:105-139)
    assert b200vfx.hash_distance(ba, bp) <= 8                     # 1 % perturbation stays close
    with pytest.raises(b200vfx.B200VfxError) as e:
        ctx.blockhash_sums("RGBA", 30, 16, a, 4 * w, np.zeros(64, np.uint32))
    assert e.value.code == b200vfx.ERR_UNSUPPORTED


@pytest.mark.parametrize("fmt,w,h", [("RGBA", 3840, 2160), ("RGBA", 72, 40), ("RGB", 640, 480), ("RGBA", 1920, 1080)])
def test_blockhash_batch_matches_per_frame(ctx, fmt, w, h):
    """videocompare hashes the reference frame and every other pad's frame per tick: one launch for up to 8 frames,
    host and device frames mixed, strides differing per frame (BASELINE config 4 = two 4K RGBA streams)"""
    torch = pytest.importorskip("torch")
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    for n in (1, 2, 3, 8):
        frames = [synth.frame_noise(fmt, w, h, 40 + i, stride=synth.default_stride(fmt, w) + (16 * (i % 2))) if i % 3 else
                  synth.frame_ramps(fmt, w, h) for i in range(n)]
        if fmt == "RGBA":
            frames[-1][::5, 3:4 * w:8] = 0
        exp = np.concatenate([orc.blockhash_sums(fmt, w, h, f) for f in frames])
        keep = [torch.from_numpy(f).cuda() if i % 2 == 0 else None for i, f in enumerate(frames)]
        srcs = [keep[i] if keep[i] is not None else frames[i] for i in range(n)]
        sums = np.full(64 * n, 0xDEADBEEF, np.uint32)
        ctx.blockhash_sums_batch(fmt, w, h, srcs, [f.shape[1] for f in frames], sums)
        assert (sums == exp).all(), n
        if n == 2:  # all-device frames, device sums: asynchronous on the context stream
            d_sums = torch.full((128,), -1, dtype=torch.int32, device="cuda")
            ctx.blockhash_sums_batch(fmt, w, h, [torch.from_numpy(f).cuda() for f in frames], [f.shape[1] for f in frames], d_sums)
            torch.cuda.synchronize()
            assert (d_sums.cpu().numpy().view(np.uint32) == exp).all()
    with pytest.raises(b200vfx.B200VfxError):
        ctx.blockhash_sums_batch(fmt, w, h, [frames[0]] * 9, [frames[0].shape[1]] * 9, np.zeros(64 * 9, np.uint32))


# ---- roundedcorners --------------------------------------------------------------------------------
@pytest.mark.parametrize("w,h,stride,r", [(64, 50, 64, 12), (1920, 1080, 1920, 64), (33, 17, 36, 5), (10, 9, 12, 100),
                                          (64, 51, 68, 0), (1, 1, 4, 3), (640, 480, 640, 1)])
def test_roundmask(ctx, w, h, stride, r):
    exp = orc.roundmask(w, h, stride, r)
    got = np.full(exp.shape, 0x77, np.uint8)
    ctx.roundmask_generate(w, h, stride, r, got)
    assert (got == exp).all()


# ---- row tiles (the multi-GPU sharding unit) -------------------------------------------------------
@pytest.mark.parametrize("world", [2, 3, 8])
def test_row_tiles_match_full_frame(ctx, world):
    """rank r calls the C ABI on (data + r0*stride, rows): tiles + seams must equal the single-call result"""
    from b200vfx import sharding
    w, h = 1280, 723
    cube = orc.cube_parse(synth.cube_text_3d(17, "mix"))
    set_cube(ctx, cube)
    frame = synth.frame_natural("RGBA", w, h, 5)
    full = gpu_colorlut(ctx, "RGBA", w, h, frame)
    tiled = np.zeros_like(frame)
    hs = frame.copy()
    for r in range(world):
        r0, r1 = sharding.row_range(h, world, r)
        if r1 > r0:
            ctx.colorlut_process("RGBA", w, r1 - r0, frame[r0:].ctypes.data, 4 * w, tiled[r0:].ctypes.data, 4 * w)
            ctx.hsvfilter_process("RGBA", w, r1 - r0, hs[r0:].ctypes.data, 4 * w, hue_shift=33.0)
    assert (tiled == full).all() and (full == orc.colorlut_apply(cube, "RGBA", w, h, frame)).all()
    assert (hs == orc.hsvfilter("RGBA", w, h, frame, hue_shift=33.0)).all()


# ---- kernel variants and PDL overlap ---------------------------------------------------------------
def test_kernel_variants_give_identical_results(ctx):
    torch = pytest.importorskip("torch")
    w, h = 1920, 270
    cube = orc.cube_parse(synth.cube_text_3d(17, "mix"))
    cube1 = orc.cube_parse(synth.cube_text_1d(256, 2.0))
    frame = synth.frame_natural("RGBA", w, h, 9)
    d_src = torch.from_numpy(frame).cuda()
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    try:
        for cb in (cube, cube1):
            set_cube(ctx, cb)
            exp = orc.colorlut_apply(cb, "RGBA", w, h, frame, threads=NT)
            for sp in (0, 1):
                for cfg in range(8):
                    for pdl in (0, 1):
                        for px in (4, 8, 16):
                            if sp == 1 and px != 4:
                                continue
                            if sp == 0 and cfg != 0:
                                continue
                            for k, v in (("stream_path", sp), ("stream_cfg", cfg), ("pdl", pdl), ("memo_px", px)):
                                ctx.set_option(k, v)
                            d_dst = torch.zeros_like(d_src)
                            for _ in range(3):
                                ctx.colorlut_process("RGBA", w, h, d_src, 4 * w, d_dst, 4 * w)
                            torch.cuda.synchronize()
                            assert (d_dst.cpu().numpy() == exp).all(), (cb.kind, sp, cfg, pdl, px)
    finally:
        for k, v in (("stream_path", -1), ("stream_cfg", 0), ("pdl", 1), ("memo_px", 4)):
            ctx.set_option(k, v)


def test_pdl_overlap_respects_dependencies(ctx):
    """back-to-back launches with PDL: chains (y = f(x); z = g(y)) across two contexts on one stream, the same
    destination written twice, and in-place work on a buffer that a previous kernel still reads"""
    torch = pytest.importorskip("torch")
    w, h = 3840, 1080
    cube_a = orc.cube_parse(synth.cube_text_3d(9, "mix"))
    cube_b = orc.cube_parse(synth.cube_text_3d(5, "invert"))
    st = torch.cuda.current_stream().cuda_stream
    ctx_b = b200vfx.Context(0)
    try:
        ctx.set_stream(st); ctx_b.set_stream(st)
        set_cube(ctx, cube_a); set_cube(ctx_b, cube_b)
        frames = [synth.frame_noise("RGBA", w, h, 40 + i) for i in range(3)]
        exp_a = [orc.colorlut_apply(cube_a, "RGBA", w, h, f, threads=NT) for f in frames]
        exp_ab = [orc.colorlut_apply(cube_b, "RGBA", w, h, e, threads=NT) for e in exp_a]
        d = [torch.from_numpy(f).cuda() for f in frames]
        y = torch.zeros_like(d[0]); z = torch.zeros_like(d[0])
        dummy_in = [torch.from_numpy(frames[0]).cuda() for _ in range(2)]
        dummy_out = [torch.zeros_like(d[0]) for _ in range(2)]
        for rep in range(3):
            for i in range(3):
                ctx.colorlut_process("RGBA", w, h, dummy_in[0], 4 * w, dummy_out[0], 4 * w)   # warm: ours, PDL-able
                ctx.colorlut_process("RGBA", w, h, dummy_in[1], 4 * w, dummy_out[1], 4 * w)
                ctx.colorlut_process("RGBA", w, h, d[i], 4 * w, y, 4 * w)                     # y = A(x)
                ctx_b.colorlut_process("RGBA", w, h, y, 4 * w, z, 4 * w)                      # z = B(y): RAW on y
                ctx.colorlut_process("RGBA", w, h, d[(i + 1) % 3], 4 * w, y, 4 * w)           # WAR on y (ctx_b still reading)
                torch.cuda.synchronize()
                assert (z.cpu().numpy() == exp_ab[i]).all()
                assert (y.cpu().numpy() == exp_a[(i + 1) % 3]).all()
        # same destination twice in a row: the second frame must win
        ctx.colorlut_process("RGBA", w, h, d[0], 4 * w, y, 4 * w)
        ctx.colorlut_process("RGBA", w, h, d[1], 4 * w, y, 4 * w)
        torch.cuda.synchronize()
        assert (y.cpu().numpy() == exp_a[1]).all()
    finally:
        ctx_b.close()


def test_hsv_memo_policy_switches_and_tracks_settings(ctx):
    """auto policy: direct evaluation until the current settings have seen 2^24 pixels, then the memo table; a
    settings change (properties are mutable in PLAYING) must never serve stale answers"""
    ctx.set_option("hsv_memo", -1)
    w, h = 2048, 1024   # 2^21 pixels per frame: the switch happens at the 8th frame
    frame = synth.frame_noise("RGBA", w, h, 21)
    det_in = synth.frame_noise("BGRx", w, h, 22)
    exp = {hs: orc.hsvfilter("RGBA", w, h, frame, hue_shift=hs, threads=NT) for hs in (10.0, 200.0)}
    dexp = {hr: orc.hsvdetector("BGRx", "RGBA", w, h, det_in, hue_ref=hr, hue_var=40.0, sat_var=1.0, val_var=1.0, threads=NT) for hr in (30.0, 250.0)}
    n0 = ctx.kernel_launches
    for hs, hr in ((10.0, 30.0), (200.0, 250.0), (10.0, 30.0)):
        for i in range(11):
            assert (gpu_hsvfilter(ctx, "RGBA", w, h, frame, hue_shift=hs) == exp[hs]).all()
            assert (gpu_hsvdetector(ctx, "BGRx", "RGBA", w, h, det_in, hue_ref=hr, hue_var=40.0, saturation_var=1.0, value_var=1.0) == dexp[hr]).all()
    assert ctx.kernel_launches - n0 >= 3 * 2 * 12   # >= 3 epochs x (11 frames + 1 table build) x 2 elements (host frames are chunked)


@pytest.mark.parametrize("fmt", ["RGBA64_LE", "RGBA64_BE"])
def test_colorlut_rgba64_every_channel_value(ctx, fmt):
    """all 65536 values in every channel: exhaustively validates the 3-instruction exact x/65535 used by the
    RGBA64 kernel (and the inline index/weight computation) against the oracle's IEEE division"""
    i = np.arange(65536, dtype=np.uint32)
    px = np.stack([i, i ^ 0x5555, 65535 - i, (i * 31) & 0xFFFF], axis=-1)
    dt = "<u2" if fmt.endswith("LE") else ">u2"
    frame = px.astype(dt).view(np.uint8).reshape(256, 256 * 8)
    for cube in (orc.cube_parse(synth.cube_text_3d(33, "mix")),
                 orc.cube_parse(synth.cube_text_3d(7, "mix", domain=((-0.3, 0.1, 0.0), (1.7, 0.8, 3.0)))),
                 orc.cube_from_values(3, 256, synth.lut_values_3d(256, "mix")) if fmt.endswith("LE") else orc.cube_parse(synth.cube_text_3d(2, "invert")),
                 orc.cube_parse(synth.cube_text_1d(65536, 1.7)),
                 orc.cube_parse(synth.cube_text_1d(3, 2.0, domain=((0.25, 0.0, -1.0), (0.75, 2.0, 1.0))))):
        set_cube(ctx, cube)
        exp = orc.colorlut_apply(cube, fmt, 256, 256, frame, threads=NT)
        assert (gpu_colorlut(ctx, fmt, 256, 256, frame) == exp).all(), (cube.kind, cube.size)


def test_colorlut_pinned_host_zero_copy_matches_staged(ctx):
    """pinned host frames take the zero-copy path (TMA bulk copies straight from/to host memory); pageable or
    mis-aligned frames take the staged copy-engine pipeline -- identical bytes either way"""
    torch = pytest.importorskip("torch")
    cube = orc.cube_parse(synth.cube_text_3d(17, "mix"))
    set_cube(ctx, cube)
    for (w, h, spad, dpad) in ((3840, 2160, 0, 0), (1280, 333, 16, 48), (1000, 64, 0, 0), (636, 50, 16, 16)):
        frame = synth.frame_natural("RGBA", w, h, 3, stride=4 * w + spad)
        exp = orc.colorlut_apply(cube, "RGBA", w, h, frame, dst_stride=4 * w + dpad, threads=NT)
        src = torch.from_numpy(frame).pin_memory()
        for zc in (1, 0):
            ctx.set_option("zero_copy", zc)
            dst = torch.full((h, 4 * w + dpad), 0x5A, dtype=torch.uint8).pin_memory()
            n0 = ctx.kernel_launches
            ctx.colorlut_process("RGBA", w, h, src.numpy(), 4 * w + spad, dst.numpy(), 4 * w + dpad)
            assert (dst.numpy() == exp).all(), (w, h, zc)
            if zc == 1 and w % 4 == 0:
                assert ctx.kernel_launches - n0 <= 2     # one kernel (+ the memo build on first use), no staging chunks
    ctx.set_option("zero_copy", 1)
