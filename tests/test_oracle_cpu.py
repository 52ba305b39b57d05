"""CPU tests: pin the oracle against the reference's own KATs, SURVEY Appendix C vectors and an
independent numpy-f32 restatement.  No GPU needed."""
import numpy as np
import pytest

import np_restatement as npr
import oracle_binding as orc
from b200vfx import synth

NT = 8


def all_colors():
    idx = np.arange(1 << 24, dtype=np.uint32)
    return (idx & 0xFF).astype(np.uint8), ((idx >> 8) & 0xFF).astype(np.uint8), ((idx >> 16) & 0xFF).astype(np.uint8)


def all_colors_frame(fmt="RGBA", alpha=None):
    """4096x4096 frame holding every 8-bit colour once (canonical r fastest)."""
    r, g, b = all_colors()
    a = (np.arange(1 << 24, dtype=np.uint32) * 7 + 3).astype(np.uint8) if alpha is None else np.full(1 << 24, alpha, np.uint8)
    planes = {"R": r, "G": g, "B": b, "A": a, "x": a}
    order = synth._ORDER[fmt]
    px = np.stack([planes[ch] for ch in order], axis=-1)
    return px.reshape(4096, 4096 * len(order))


# ---- .cube parser: the reference's own unit tests (video/colorlut/src/parser.rs:382-473) -------
def test_parse_3d_lut():
    text = """
            LUT_3D_SIZE 2

            0.0 0.0 0.0
            1.0 0.0 0.0
            0.0 1.0 0.0
            1.0 1.0 0.0
            0.0 0.0 1.0
            1.0 0.0 1.0
            0.0 1.0 1.0
            1.0 1.0 1.0
        """
    lut = orc.cube_parse(text)
    assert lut.kind == 3 and lut.size == 2 and lut.values.shape == (8, 3)
    assert lut.values[0].tolist() == [0.0, 0.0, 0.0]          # at(0,0,0)
    assert lut.values[1 + 2 + 4].tolist() == [1.0, 1.0, 1.0]  # at(1,1,1)


def test_keyword_after_lut_size():
    text = """
            LUT_1D_SIZE 2

            TITLE "test"
            DOMAIN_MIN 0.0 0.0 0.0
            DOMAIN_MAX 1.0 1.0 1.0

            0.0 0.0 0.0
            1.0 0.5 0.7
        """
    lut = orc.cube_parse(text)
    assert lut.kind == 1 and lut.size == 2
    assert lut.values[:, 0].tolist() == [0.0, 1.0]
    assert lut.values[:, 1].tolist() == [0.0, 0.5]
    assert lut.values[:, 2].tolist() == [0.0, np.float32(0.7)]


@pytest.mark.parametrize("text", [
    "LUT_1D_SIZE 2\n\n0.0 0.0 0.0\n1.0 0.0 0.0\nTITLE \"invalid\"\n",      # keyword_after_data
    "LUT_1D_SIZE 2\n\n0.0 0.0 0.0\nTITLE \"invalid\"\n1.0 0.0 0.0\n",      # keyword_between_data
    "LUT_1D_SIZE 2\nLUT_3D_SIZE 2\n\n0.0 0.0 0.0\n1.0 1.0 1.0\n",          # multiple_lut_sizes
])
def test_parser_reference_error_cases(text):
    with pytest.raises(orc.CubeError) as e:
        orc.cube_parse(text)
    assert e.value.code == -1


@pytest.mark.parametrize("text", [
    "0 0 0\nLUT_1D_SIZE 2\n",                       # data before size
    "LUT_1D_SIZE 2\n0 0 0\n",                       # count mismatch
    "LUT_3D_SIZE 2\n" + "0 0 0\n" * 7,              # count mismatch 3D
    "LUT_1D_SIZE 1\n0 0 0\n",                       # size < 2
    "LUT_3D_SIZE 257\n",                            # size > 256
    "LUT_1D_SIZE 65537\n",                          # size > 65536
    "LUT_1D_SIZE 2\n0 0\n1 1 1\n",                  # too few floats
    "LUT_1D_SIZE 2\n0 0 0 0\n1 1 1\n",              # too many floats
    "LUT_1D_SIZE 2\n0 0 0x1p0\n1 1 1\n",            # hex float rejected by Rust
    "LUT_1D_SIZE 2\n0 0 1f\n1 1 1\n",               # suffix rejected
    "LUT_1D_SIZE -2\n",                             # usize parse
    "LUT_1D_SIZE 2 3\n",                            # extra token
    "LUT_1D_SIZE 2\nDOMAIN_MIN 1 0 0\nDOMAIN_MAX 1 1 1\n0 0 0\n1 1 1\n",  # min >= max
    "LUT_1D_SIZE 2\nLUT_3D_INPUT_RANGE 0 1\n0 0 0\n1 1 1\n",   # unknown keyword = bad float
    "﻿LUT_1D_SIZE 2\n0 0 0\n1 1 1\n",          # BOM: first token is not a keyword
    "",                                             # missing size
    "# only a comment\n",
])
def test_parser_more_error_cases(text):
    with pytest.raises(orc.CubeError) as e:
        orc.cube_parse(text)
    assert e.value.code == -1


def test_parser_invalid_utf8_is_io_error():
    with pytest.raises(orc.CubeError) as e:
        orc.cube_parse(b"LUT_1D_SIZE 2\n0 0 0\n1 1 \xff1\n")
    assert e.value.code == -2


def test_parser_accepts_rust_float_forms_and_domain():
    text = ("# c\r\nTITLE \"x y z\"\r\nLUT_1D_SIZE +3\r\nDOMAIN_MIN -1 0 .5\nDOMAIN_MAX 1 2 1.\n"
            "1e-1 +.5 5.E-1\n inf -Infinity NaN \n\t0.1 0.2　0.3\n")
    lut = orc.cube_parse(text)
    assert lut.kind == 1 and lut.size == 3
    assert lut.values[0].tolist() == [np.float32(0.1), 0.5, 0.5]
    assert np.isposinf(lut.values[1, 0]) and np.isneginf(lut.values[1, 1]) and np.isnan(lut.values[1, 2])
    assert lut.values[2].tolist() == [np.float32(0.1), np.float32(0.2), np.float32(0.3)]
    f = np.float32
    exp_scale = [f(1) / (f(1) - f(-1)), f(1) / (f(2) - f(0)), f(1) / (f(1) - f(0.5))]
    assert lut.scale.tolist() == exp_scale
    assert lut.offset.tolist() == [-f(-1) * exp_scale[0], -f(0) * exp_scale[1], -f(0.5) * exp_scale[2]]
    d = orc.cube_parse("LUT_1D_SIZE 2\n0 0 0\n1 1 1")
    assert d.scale.tolist() == [1, 1, 1] and all(np.signbit(d.offset))  # -0.0 * 1.0 = -0.0


def test_generated_cube_roundtrip():
    lut = orc.cube_parse(synth.cube_text_3d(5, "mix", title="mix"))
    assert lut.kind == 3 and lut.size == 5 and lut.values.shape == (125, 3)
    assert lut.values[4].tolist() == [1.0, 0.0, np.float32("%.6f" % (4 / 12))]


# ---- hsvutils: the reference's own unit tests (video/hsv/src/hsvutils.rs:237-279) --------------
HSV_KATS = [((255, 255, 255), (0.0, 0.0, 1.0)), ((0, 0, 0), (0.0, 0.0, 0.0)), ((255, 0, 0), (0.0, 1.0, 1.0)),
            ((0, 255, 0), (120.0, 1.0, 1.0)), ((0, 0, 255), (240.0, 1.0, 1.0))]


def _equiv(hsv, exp, eps=1e-5):
    f = np.float32
    sh = f(hsv[0]) + (f(180) - f(exp[0]))
    if sh < 0:
        sh += f(360)
    sh = np.fmod(f(sh), f(360))
    return abs(sh - 180) < eps and abs(hsv[1] - exp[1]) < eps and abs(hsv[2] - exp[2]) < eps


@pytest.mark.parametrize("rgb,hsv", HSV_KATS)
def test_hsvutils_kats(rgb, hsv):
    assert _equiv(orc.hsv_from(rgb), hsv)
    assert _equiv(orc.hsv_from(rgb[::-1], bgr=True), hsv)
    assert orc.hsv_to(hsv) == list(rgb)
    assert orc.hsv_to(hsv, bgr=True) == list(rgb[::-1])


# ---- SURVEY Appendix C derived vectors + numpy cross-check over all 2^24 colours ---------------
def test_hsvfilter_defaults_all_colors():
    frame = all_colors_frame("RGBA")
    out = orc.hsvfilter("RGBA", 4096, 4096, frame, threads=NT)
    i = frame.reshape(-1, 4).astype(np.int16)
    o = out.reshape(-1, 4).astype(np.int16)
    assert (o[:, 3] == i[:, 3]).all()
    delta = o[:, :3] - i[:, :3]
    assert int((delta != 0).any(axis=1).sum()) == 11093274
    assert set(np.unique(delta).tolist()) <= {-1, 0}
    r, g, b = all_colors()
    nr = npr.hsvfilter_rgb(r, g, b)
    for k in range(3):
        assert (o[:, k] == nr[k]).all()


@pytest.mark.parametrize("rgb,kw,exp", [
    ((0, 1, 3), {}, (0, 0, 3)), ((0, 1, 7), {}, (0, 0, 7)), ((1, 2, 3), {}, (1, 2, 3)),
    ((254, 255, 253), dict(hue_shift=359.9), (254, 255, 253)),
    ((255, 0, 0), dict(hue_shift=120.0), (0, 255, 0)),
    ((10, 200, 30), dict(hue_shift=-90.5), (200, 123, 9)),
    ((100, 50, 25), dict(hue_shift=45.0, sat_mul=0.5, val_off=0.1), (121, 125, 78)),
])
def test_hsvfilter_vectors(rgb, kw, exp):
    px = np.array([[rgb[0], rgb[1], rgb[2], 9]], np.uint8)
    assert orc.hsvfilter("RGBA", 1, 1, px, **kw)[0].tolist() == list(exp) + [9]
    bgr = np.array([[rgb[2], rgb[1], rgb[0]] + [0]], np.uint8)  # 3-bpp row padded to 4
    assert orc.hsvfilter("BGR", 1, 1, bgr, **kw)[0, :3].tolist() == list(exp)[::-1]


def test_hsvfilter_hsv_intermediates():
    f = np.float32
    assert orc.hsv_from((10, 200, 30)) == [f(126.3157958984375), f(0.949999988079071), f(0.7843137383460999)]
    assert orc.hsv_from((100, 50, 25)) == [f(19.999998092651367), f(0.75), f(0.3921568691730499)]


@pytest.mark.parametrize("kw", [dict(hue_shift=90.0), dict(hue_shift=-270.25, sat_mul=1.7, sat_off=-0.2, val_mul=0.8, val_off=0.15),
                                dict(hue_shift=1e9), dict(sat_mul=float("nan")), dict(val_off=float("inf"), hue_shift=float("-inf"))])
def test_hsvfilter_matches_numpy_all_colors(kw):
    frame = all_colors_frame("xBGR")
    r, g, b = all_colors()
    out = orc.hsvfilter("xBGR", 4096, 4096, frame, threads=NT, **kw).reshape(-1, 4)
    with np.errstate(invalid="ignore"):
        nr = npr.hsvfilter_rgb(r, g, b, **kw)
    assert (out[:, 0] == frame.reshape(-1, 4)[:, 0]).all()
    assert (out[:, 3] == nr[0]).all() and (out[:, 2] == nr[1]).all() and (out[:, 1] == nr[2]).all()


def test_hsvdetector_defaults_all_colors():
    frame = all_colors_frame("RGBx")
    out = orc.hsvdetector("RGBx", "RGBA", 4096, 4096, frame, threads=NT).reshape(-1, 4)
    assert (out[:, :3] == frame.reshape(-1, 4)[:, :3]).all()
    assert set(np.unique(out[:, 3]).tolist()) == {0, 255}
    assert int((out[:, 3] == 255).sum()) == 719
    col = lambda r, g, b: out[r | (g << 8) | (b << 16), 3]
    assert col(0, 0, 0) == 255 and col(76, 76, 76) == 255 and col(77, 77, 77) == 0
    r, g, b = all_colors()
    assert ((out[:, 3] == 255) == npr.hsvdetect_rgb(r, g, b)).all()


def test_hsvdetector_config3_all_colors():
    kw = dict(hue_ref=120.0, hue_var=30.0, sat_ref=0.8, sat_var=0.2, val_ref=0.8, val_var=0.2)
    frame = all_colors_frame("BGRx")
    out = orc.hsvdetector("BGRx", "ARGB", 4096, 4096, frame, threads=NT, **kw).reshape(-1, 4)
    fr = frame.reshape(-1, 4)
    assert (out[:, 1] == fr[:, 2]).all() and (out[:, 2] == fr[:, 1]).all() and (out[:, 3] == fr[:, 0]).all()
    assert int((out[:, 0] == 255).sum()) == 1415062
    col = lambda r, g, b: out[r | (g << 8) | (b << 16), 0]
    assert col(10, 200, 30) == 255 and col(0, 255, 0) == 255 and col(30, 200, 10) == 255
    r, g, b = all_colors()
    assert ((out[:, 0] == 255) == npr.hsvdetect_rgb(r, g, b, **kw)).all()


IN_FMTS = ["RGBx", "xRGB", "BGRx", "xBGR", "RGB", "BGR"]
OUT_FMTS = ["RGBA", "ARGB", "BGRA", "ABGR"]


@pytest.mark.parametrize("ifmt", IN_FMTS)
@pytest.mark.parametrize("ofmt", OUT_FMTS)
def test_hsvdetector_format_matrix(ifmt, ofmt):
    w, h = 37, 5
    src = synth.frame_noise(ifmt, w, h, 0x5EED0003, stride=synth.default_stride(ifmt, w) + 8)
    out = orc.hsvdetector(ifmt, ofmt, w, h, src, dst_stride=4 * w + 12, hue_ref=200.0, hue_var=90.0, sat_ref=0.5,
                          sat_var=0.5, val_ref=0.5, val_var=0.5)
    bpp = 3 if ifmt in ("RGB", "BGR") else 4
    ioff = 1 if ifmt in ("xRGB", "xBGR") else 0
    px = src[:, :bpp * w].reshape(h, w, bpp)[:, :, ioff:ioff + 3]
    rgb = px[:, :, ::-1] if "BGR" in ifmt else px
    hit = npr.hsvdetect_rgb(rgb[..., 0], rgb[..., 1], rgb[..., 2], hue_ref=200.0, hue_var=90.0, sat_ref=0.5,
                            sat_var=0.5, val_ref=0.5, val_var=0.5)
    o = out[:, :4 * w].reshape(h, w, 4)
    ooff = 1 if ofmt in ("ARGB", "ABGR") else 0
    ocol = o[:, :, ooff:ooff + 3]
    orgb = ocol[:, :, ::-1] if "BGR" in ofmt else ocol
    assert (orgb == rgb).all()
    assert (o[:, :, 0 if ooff else 3] == np.where(hit, 255, 0)).all()
    assert (out[:, 4 * w:] == 0x5A).all()  # row padding untouched


def _cube(n, kind):
    return orc.cube_from_values(3, n, synth.lut_values_3d(n, kind))


@pytest.mark.parametrize("n", [2, 33, 65])
def test_colorlut_identity_and_invert_exact(n):
    frame = all_colors_frame("RGBA")
    out = orc.colorlut_apply(_cube(n, "identity"), "RGBA", 4096, 4096, frame, threads=NT)
    assert (out == frame).all()
    out = orc.colorlut_apply(_cube(n, "invert"), "RGBA", 4096, 4096, frame, threads=NT).reshape(-1, 4)
    fr = frame.reshape(-1, 4)
    assert (out[:, :3] == 255 - fr[:, :3]).all() and (out[:, 3] == fr[:, 3]).all()


@pytest.mark.parametrize("n,exp", [
    (33, [(35, 182, 140), (185, 230, 220), (1, 204, 70)]),
    (65, [(35, 182, 140), (185, 230, 220), (1, 204, 70)]),
    (2, [(95, 130, 140), (217, 207, 220), (15, 163, 70)]),
])
def test_colorlut_mix_vectors(n, exp):
    px = np.array([[95, 130, 194, 1, 217, 207, 235, 2, 15, 163, 33, 3]], np.uint8)
    out = orc.colorlut_apply(_cube(n, "mix"), "RGBA", 3, 1, px)[0].reshape(3, 4)
    assert [tuple(p[:3]) for p in out.tolist()] == exp and out[:, 3].tolist() == [1, 2, 3]


def test_colorlut_1d_vectors():
    n = 1024
    tab = (np.arange(n, dtype=np.float32) / np.float32(1023))
    tab = (tab * tab).astype(np.float32)
    cube = orc.cube_from_values(1, n, np.stack([tab] * 3, axis=-1))
    px = np.array([[0, 1, 127, 9, 128, 254, 255, 8]], np.uint8)
    out = orc.colorlut_apply(cube, "RGBA", 2, 1, px)[0].tolist()
    assert out == [0, 0, 63, 9, 64, 253, 255, 8]
    for fmt, dt in (("RGBA64_LE", "<u2"), ("RGBA64_BE", ">u2")):
        px16 = np.array([[0, 1, 32767, 0x1234, 32768, 65534, 65535, 0xBEEF]], dt).view(np.uint8)
        o = orc.colorlut_apply(cube, fmt, 2, 1, px16).view(dt)[0].tolist()
        assert o == [0, 0, 16383, 0x1234, 16384, 65533, 65535, 0xBEEF]


@pytest.mark.parametrize("n,kind,dom", [(33, "mix", None), (17, "mix", ((-0.25, 0.0, 0.1), (1.5, 0.75, 0.9))),
                                        (65, "mix", None), (2, "invert", None)])
def test_colorlut_3d_matches_numpy_all_colors(n, kind, dom):
    cube = orc.cube_parse(synth.cube_text_3d(n, kind, domain=dom))
    frame = all_colors_frame("RGBA")
    out = orc.colorlut_apply(cube, "RGBA", 4096, 4096, frame, threads=NT).reshape(-1, 4)
    r, g, b = all_colors()
    nr = npr.colorlut_3d(cube.values, n, cube.scale, cube.offset, r, g, b)
    for k in range(3):
        assert (out[:, k] == nr[k]).all()


def test_colorlut_u16_matches_numpy():
    cube = orc.cube_parse(synth.cube_text_3d(33, "mix", domain=((0.0, 0.1, 0.0), (1.0, 0.9, 2.0))))
    w, h = 1024, 256
    for fmt, dt in (("RGBA64_LE", "<u2"), ("RGBA64_BE", ">u2")):
        src = synth.frame_noise(fmt, w, h, 77, stride=8 * w + 16)
        out = orc.colorlut_apply(cube, fmt, w, h, src, dst_stride=8 * w + 32, threads=NT)
        s16 = src[:, :8 * w].copy().view(dt).reshape(h, w, 4)
        o16 = out[:, :8 * w].copy().view(dt).reshape(h, w, 4)
        nr = npr.colorlut_3d(cube.values, 33, cube.scale, cube.offset, s16[..., 0].astype(np.int64),
                             s16[..., 1].astype(np.int64), s16[..., 2].astype(np.int64), denom=65535, outmax=65535)
        for k in range(3):
            assert (o16[..., k] == nr[k]).all()
        assert (o16[..., 3] == s16[..., 3]).all()
        assert (out[:, 8 * w:] == 0x5A).all()
    cube1 = orc.cube_parse(synth.cube_text_1d(4096, 2.2, domain=((0.0, 0.0, 0.0), (0.5, 1.0, 2.0))))
    src = synth.frame_noise("RGBA64_LE", w, h, 78)
    out = orc.colorlut_apply(cube1, "RGBA64_LE", w, h, src, threads=NT).view("<u2").reshape(h, w, 4)
    s16 = src.view("<u2").reshape(h, w, 4)
    for k in range(3):
        e = npr.colorlut_1d(cube1.values, 4096, cube1.scale, cube1.offset, s16[..., k].astype(np.int64), k, 65535, 65535)
        assert (out[..., k] == e).all()


def test_colorlut_nan_inf_lut_entries():
    v = synth.lut_values_3d(5, "mix")
    v[7] = [np.nan, np.inf, -np.inf]
    v[60] = [1e30, -1e30, np.nan]
    cube = orc.cube_from_values(3, 5, v)
    frame = all_colors_frame("RGBA")[:256]
    out = orc.colorlut_apply(cube, "RGBA", 4096, 256, frame, threads=NT).reshape(-1, 4)
    fr = frame.reshape(-1, 4)
    with np.errstate(invalid="ignore", over="ignore"):
        nr = npr.colorlut_3d(cube.values, 5, cube.scale, cube.offset, fr[:, 0], fr[:, 1], fr[:, 2])
    for k in range(3):
        assert (out[:, k] == nr[k]).all()


# ---- blockhash / roundmask ----------------------------------------------------------------------
def test_blockhash_sums_and_distance():
    w, h = 64, 48
    a = synth.frame_noise("RGBA", w, h, 0x5EED0004, stride=4 * w + 4)
    sums = orc.blockhash_sums("RGBA", w, h, a)
    px = a[:, :4 * w].reshape(h, w, 4).astype(np.uint32)
    s = np.where(px[..., 3] == 0, 765, px[..., :3].sum(-1))
    exp = s.reshape(8, h // 8, 8, w // 8).sum(axis=(1, 3)).reshape(-1)
    assert (sums == exp).all()
    rgb = synth.frame_noise("RGB", w, h, 5)
    sums3 = orc.blockhash_sums("RGB", w, h, rgb)
    p3 = rgb[:, :3 * w].reshape(h, w, 3).astype(np.uint32).sum(-1)
    assert (sums3 == p3.reshape(8, h // 8, 8, w // 8).sum(axis=(1, 3)).reshape(-1)).all()
    # identical solid frames -> all bits 0 -> distance 0 (tests/videocompare.rs:57-103)
    red = synth.frame_solid("RGBA", w, h)
    b1 = orc.blockhash_bits(orc.blockhash_sums("RGBA", w, h, red), w, h)
    assert b1.sum() == 0
    b2 = orc.blockhash_bits(sums, w, h)
    assert int((b1 != b2).sum()) > 0  # snow vs red -> distance > 0 (:105-139)


def test_roundmask_properties():
    w, h, stride, r = 64, 50, 64, 12
    m = orc.roundmask(w, h, stride, r)
    assert m.shape == (50, 64)
    assert (m[r:h - r, :] == 255).all() and (m[:, r:w - r][:h] == 255).all()
    assert m[0, 0] == 0 and m[0, w - 1] == 0 and m[h - 1, 0] == 0 and m[h - 1, w - 1] == 0
    assert (m[:h, :w] == m[:h, :w][::-1, :]).all() and (m[:h, :w] == m[:h, :w][:, ::-1]).all()
    yy, xx = np.mgrid[0:r, 0:r]
    dist = np.hypot(r - xx - 0.5, r - yy - 0.5)
    assert (m[:r, :r][dist < r - 1.3] == 255).all() and (m[:r, :r][dist > r + 1.3] == 0).all()
    assert (orc.roundmask(w, 51, 68, 0) == 255).all()
    m2 = orc.roundmask(w, 51, 68, 5)
    assert m2.shape == (52, 68) and (m2[51] == 0).all() and (m2[:, 64:] == 0).all()
