"""colordetect (video/videofx/src/colordetect/imp.rs, video/videofx/tests/colordetect.rs).

The per-pixel pass (the 5-bit histogram of color-thief's get_palette over every `quality`-th pixel of the flat plane)
is the GPU kernel and is compared bit-for-bit with the oracle loop; the median cut and the CSS colour name are host
work whose only reference-pinned behaviour is "a red frame names red" (third-party crates, parity otherwise unpinned):
there the product's C++ and the oracle's C restatement are compared with each other."""
import numpy as np
import pytest

import b200vfx
import oracle_binding as orc
from b200vfx import gst, synth

FORMATS = {"RGB": (3, 0, 1, 2, None), "RGBA": (4, 0, 1, 2, 3), "ARGB": (4, 1, 2, 3, 0), "BGR": (3, 2, 1, 0, None), "BGRA": (4, 2, 1, 0, 3)}


def np_histogram(fmt, plane, quality):
    """independent numpy restatement of make_histogram_and_vbox's pixel loop"""
    bpp, ro, go, bo, ao = FORMATS[fmt]
    flat = plane.reshape(-1)
    n = flat.size // bpp
    px = flat[:n * bpp].reshape(n, bpp)[::quality].astype(np.uint32)
    r, g, b = px[:, ro], px[:, go], px[:, bo]
    a = px[:, ao] if ao is not None else np.full(len(px), 255, np.uint32)
    keep = (a >= 125) & ~((r > 250) & (g > 250) & (b > 250))
    idx = ((r >> 3) << 10) + ((g >> 3) << 5) + (b >> 3)
    return np.bincount(idx[keep], minlength=32768).astype(np.uint32)


def plane_for(fmt, w, h, seed, pad=0, kind="noise"):
    bpp = FORMATS[fmt][0]
    rng = np.random.default_rng(seed)
    p = np.zeros((h, bpp * w + pad), np.uint8)
    if kind == "noise":
        p[:] = rng.integers(0, 256, p.shape, dtype=np.uint8)       # padding bytes are sampled too: fill them
    else:  # smooth gradients with a little noise, some transparent and some white pixels
        yy, xx = np.mgrid[0:h, 0:p.shape[1]]
        p[:] = ((xx // bpp) * 255 // max(w - 1, 1) + (yy * 3) + rng.integers(0, 4, p.shape)).astype(np.uint8)
        p[::7, ::5] = 255
        p[1::9, 3::11] = 0
    return p


# ---- CPU: oracle and host logic ----------------------------------------------------------------------
@pytest.mark.parametrize("fmt", list(FORMATS))
@pytest.mark.parametrize("quality", [1, 3, 10])
def test_oracle_histogram_matches_numpy(fmt, quality):
    for (w, h, pad, kind) in ((64, 16, 0, "noise"), (61, 13, 5, "smooth"), (1, 1, 0, "noise"), (33, 7, 3, "noise")):
        plane = plane_for(fmt, w, h, 11 + w, pad, kind)
        assert (orc.colordetect_histogram(fmt, w, h, plane, quality) == np_histogram(fmt, plane, quality)).all()


def test_red_frame_names_red_like_the_reference_test():
    """video/videofx/tests/colordetect.rs:21-71: videotestsrc pattern=red -> dominant-color "red" """
    w, h = 320, 240
    red = synth.frame_solid("RGBA", w, h)
    hist = orc.colordetect_histogram("RGBA", w, h, red, 10)
    assert hist[31 << 10] == (w * h + 9) // 10 and hist.sum() == hist[31 << 10]
    for pal in (orc.colordetect_palette(hist, 2), b200vfx.colordetect_palette(hist, 2)):
        assert pal[0] == (252, 4, 4)
        assert orc.css_similar(*pal[0]) == "red" and b200vfx.css_color_similar(*pal[0]) == "red"


def test_palette_product_matches_oracle_restatement():
    rng = np.random.default_rng(5)
    cases = []
    for k in range(12):
        h = np.zeros(32768, np.uint32)
        n = int(rng.integers(1, 400))
        idx = rng.integers(0, 32768, n)
        h[idx] = rng.integers(1, 5000, n)
        cases.append(h)
    cases.append(np.zeros(32768, np.uint32))                           # nothing counted (e.g. an all-white frame)
    one = np.zeros(32768, np.uint32); one[12345] = 1; cases.append(one)  # a single pixel cannot be cut
    dense = rng.integers(0, 50, 32768).astype(np.uint32); cases.append(dense)
    for plane_kind in ("noise", "smooth"):
        p = plane_for("RGBA", 256, 128, 3, 0, plane_kind)
        cases.append(orc.colordetect_histogram("RGBA", 256, 128, p, 1))
    for h in cases:
        for mc in (2, 3, 5, 8, 16, 255):
            a, b = orc.colordetect_palette(h, mc), b200vfx.colordetect_palette(h, mc)
            assert a == b and len(a) >= 1
    assert b200vfx.colordetect_palette(np.zeros(32768, np.uint32), 2)[0] == (255, 255, 255)
    with pytest.raises(b200vfx.B200VfxError):
        b200vfx.colordetect_palette(cases[0], 1)


def test_css_names_product_matches_oracle():
    rng = np.random.default_rng(9)
    for r, g, b in rng.integers(0, 256, (3000, 3)):
        assert b200vfx.css_color_similar(r, g, b) == orc.css_similar(r, g, b)
    assert b200vfx.css_color_similar(0, 255, 255) == "aqua"          # first of two names for one RGB
    assert b200vfx.css_color_similar(0, 0, 0) == "black" and b200vfx.css_color_similar(254, 254, 254) == "white"
    assert b200vfx.css_color_similar(4, 132, 4) == "green" and b200vfx.css_color_similar(4, 4, 252) == "blue"


def test_colordetect_element_surface():
    el = gst.Element("colordetect")
    assert (el.type_name, el.plugin_name) == ("GstColorDetect", "rsvideofx")
    props = {p[0]: p[1:] for p in el.list_properties()}
    assert props["quality"] == ("guint", "10", "0", "10", "mutable-playing")          # imp.rs:126-133
    assert props["max-colors"] == ("guint", "2", "2", "255", "mutable-playing")       # imp.rs:134-141
    assert el.set_property("quality", 11) != 0 and el.set_property("max-colors", 1) != 0
    assert el.pad_template_formats(gst.PAD_SINK) == ["RGB", "RGBA", "ARGB", "BGR", "BGRA"]   # imp.rs:214-222
    assert el.transform_caps(gst.PAD_SINK, ["BGRA", "I420"]) == ["BGRA"]
    f = gst.frame("RGBA", 4, 4, [np.zeros((4, 16), np.uint8)], [16])
    assert el.transform_frame_ip(f) == gst.FLOW_NOT_NEGOTIATED and "no state" in el.last_error.lower()   # imp.rs:62-65


# ---- GPU: the histogram kernel through the C ABI --------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("fmt", list(FORMATS))
def test_gpu_histogram_matches_oracle(fmt):
    with b200vfx.Context(0) as ctx:
        for (w, h, pad, kind) in ((640, 480, 0, "smooth"), (641, 37, 0, "noise"), (333, 21, 7, "smooth"), (1, 1, 0, "noise"),
                                  (1920, 64, 0, "noise"), (35, 3, 1, "noise")):
            plane = plane_for(fmt, w, h, 100 + w, pad, kind)
            for quality in (1, 2, 7, 10):
                got = np.full(32768, 0xDEAD, np.uint32)
                ctx.colordetect_histogram(fmt, w, h, plane, plane.shape[1], quality, got)
                assert (got == orc.colordetect_histogram(fmt, w, h, plane, quality)).all(), (w, h, pad, quality)


@pytest.mark.gpu
def test_gpu_histogram_device_pointers_unaligned_and_errors():
    torch = pytest.importorskip("torch")
    w, h = 500, 40
    with b200vfx.Context(0) as ctx:
        for fmt in ("RGBA", "BGR"):
            bpp = FORMATS[fmt][0]
            plane = plane_for(fmt, w, h, 3, 0, "smooth")
            big = torch.zeros(plane.size + 64, dtype=torch.uint8, device="cuda")
            for off in (0, 1, 2, 16):                                  # unaligned device base
                big[off:off + plane.size] = torch.from_numpy(plane.reshape(-1)).cuda()
                d_hist = torch.full((32768,), 7, dtype=torch.int32, device="cuda")
                ctx.colordetect_histogram(fmt, w, h, big.data_ptr() + off, bpp * w, 3, d_hist)
                ctx.synchronize()
                assert (d_hist.cpu().numpy().view(np.uint32) == orc.colordetect_histogram(fmt, w, h, plane, 3)).all()
        got = np.ones(32768, np.uint32)
        ctx.colordetect_histogram("RGBA", 0, 0, None, 0, 10, got)      # empty frame: all-zero histogram
        assert not got.any()
        for bad_q in (0, 11):
            with pytest.raises(b200vfx.B200VfxError):
                ctx.colordetect_histogram("RGBA", w, h, plane_for("RGBA", w, h, 1), 4 * w, bad_q, got)
        with pytest.raises(b200vfx.B200VfxError):
            ctx.colordetect_histogram("BGRx", w, h, plane_for("RGBA", w, h, 1), 4 * w, 10, got)


@pytest.mark.gpu
@pytest.mark.parametrize("quality", [1, 10])
def test_gpu_histogram_back_to_back_launches_overlap_safely(quality):
    """consecutive histogram launches overlap (the plane is read while the previous launch drains; the global histogram and
    the counters are touched after griddepcontrol.wait): a long unsynchronised train over few histograms, and a plane that
    the launch just before produced, stay exact"""
    torch = pytest.importorskip("torch")
    w, h = 1920, 1080
    with b200vfx.Context(0) as ctx:
        ctx.set_stream(torch.cuda.current_stream().cuda_stream)
        frames = [synth.frame_noise("RGBA", w, h, 70 + i) if i % 2 else synth.frame_natural("RGBA", w, h, 70 + i) for i in range(4)]
        exp = [orc.colordetect_histogram("RGBA", w, h, f, quality) for f in frames]
        dev = [torch.from_numpy(f).cuda() for f in frames]
        hists = [torch.full((32768,), 9, dtype=torch.int32, device="cuda") for _ in range(3)]
        for i in range(40):
            ctx.colordetect_histogram("RGBA", w, h, dev[i % 4], 4 * w, quality, hists[i % 3])
        torch.cuda.synchronize()
        for k in range(3):
            last = max(i for i in range(40) if i % 3 == k)
            assert (hists[k].cpu().numpy().view(np.uint32) == exp[last % 4]).all(), k
        cube = orc.cube_parse(synth.cube_text_3d(9, "mix"))
        ctx.colorlut_set_lut(cube.kind, cube.size, cube.values, cube.scale, cube.offset)
        want = [orc.colordetect_histogram("RGBA", w, h, orc.colorlut_apply(cube, "RGBA", w, h, f), quality) for f in frames]
        out = torch.empty_like(dev[0])
        got = [torch.zeros(32768, dtype=torch.int32, device="cuda") for _ in range(12)]
        for i in range(12):   # colorlut rewrites the plane the previous histogram launch is reading
            ctx.colorlut_process("RGBA", w, h, dev[i % 4], 4 * w, out, 4 * w)
            ctx.colordetect_histogram("RGBA", w, h, out, 4 * w, quality, got[i])
        torch.cuda.synchronize()
        for i in range(12):
            assert (got[i].cpu().numpy().view(np.uint32) == want[i % 4]).all(), i


@pytest.mark.gpu
@pytest.mark.parametrize("quality", [1, 10])
def test_gpu_histogram_4k_full_size(quality):
    """3840x2160 RGBA (the BASELINE frame shape): oracle equality plus the size-independent checksum
    sum(hist) == number of sampled pixels that pass the alpha / white test"""
    w, h = 3840, 2160
    with b200vfx.Context(0) as ctx:
        for frame in (synth.frame_ramps("RGBA", w, h), synth.frame_noise("RGBA", w, h, 0x5EED0002)):
            got = np.zeros(32768, np.uint32)
            ctx.colordetect_histogram("RGBA", w, h, frame, 4 * w, quality, got)
            px = frame.reshape(-1, 4)[::quality]
            counted = (px[:, 3] >= 125) & ~((px[:, 0] > 250) & (px[:, 1] > 250) & (px[:, 2] > 250))
            assert int(got.sum()) == int(counted.sum())
            assert (got == orc.colordetect_histogram("RGBA", w, h, frame, quality)).all()


@pytest.mark.gpu
def test_colordetect_element_red_then_blue():
    """tests/colordetect.rs: two red frames -> exactly ONE message naming "red"; a colour change posts another"""
    w, h = 320, 240
    el = gst.Element("colordetect")
    assert el.start() == 0
    assert el.set_caps("RGBA", "RGBA", w, h) == 0 and el.passthrough
    red = synth.frame_solid("RGBA", w, h)
    keep = red.copy()
    fr = lambda a: gst.frame("RGBA", w, h, [a], [4 * w])
    assert el.transform_frame_ip(fr(red)) == gst.FLOW_OK
    msg = el.pop_message()
    assert msg is not None and msg.startswith("colordetect, dominant-color=(string)red, palette=(uint){ %d" % ((252 << 16) | (4 << 8) | 4))
    assert el.transform_frame_ip(fr(red)) == gst.FLOW_OK and el.pop_message() is None
    assert (red == keep).all()                                             # passthrough: the frame is not modified
    blue = red.copy(); blue[:, 0::4] = 0; blue[:, 2::4] = 255
    assert el.transform_frame_ip(fr(blue)) == gst.FLOW_OK
    assert "dominant-color=(string)blue" in el.pop_message()
    el.set_property("quality", 0)                                           # allowed by the property, rejected by color-thief
    assert el.transform_frame_ip(fr(blue)) == gst.FLOW_ERROR
    el.set_property("quality", 1); el.set_property("max-colors", 5)
    noise = synth.frame_noise("RGBA", w, h, 4)
    assert el.transform_frame_ip(fr(noise)) == gst.FLOW_OK
    msg = el.pop_message()
    hist = orc.colordetect_histogram("RGBA", w, h, noise, 1)
    pal = orc.colordetect_palette(hist, 5)
    assert msg is not None and ("dominant-color=(string)%s," % orc.css_similar(*pal[0])) in msg
    assert msg.count(",") == len(pal) + 1                                  # n-1 commas inside the list + 2 field separators
    el.stop()
