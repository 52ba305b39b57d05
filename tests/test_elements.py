"""Element-level tests in the style of the reference's GStreamer tests (video/videofx/tests/videocompare.rs,
gst_check::Harness usage elsewhere): make the element by factory name, set properties, negotiate, push frames.
The property / caps parts run on CPU; everything that touches pixels is marked gpu."""
import numpy as np
import pytest

import oracle_binding as orc
from b200vfx import gst, synth


# ---- the drop-in surface -------------------------------------------------------
def test_factories_types_and_plugins():
    exp = {"colorlut": ("GstColorLut", "colorlut"), "hsvfilter": ("GstHsvFilter", "hsv"), "hsvdetector": ("GstHsvDetector", "hsv"),
           "roundedcorners": ("GstRoundedCorners", "rsvideofx"), "videocompare": ("GstVideoCompare", "rsvideofx")}
    for f, (t, p) in exp.items():
        el = gst.Element(f)
        assert (el.type_name, el.plugin_name) == (t, p)
    with pytest.raises(ValueError):
        gst.Element("d3d12colorlut")


def test_property_tables_match_reference():
    props = {f: {p[0]: p[1:] for p in gst.Element(f).list_properties()} for f in
             ("colorlut", "hsvfilter", "hsvdetector", "roundedcorners", "videocompare")}
    assert props["colorlut"] == {"location": ("gchararray", "NULL", "", "", "mutable-ready")}
    assert set(props["hsvfilter"]) == {"hue-shift", "saturation-mul", "saturation-off", "value-mul", "value-off"}
    assert props["hsvfilter"]["saturation-mul"][:2] == ("gfloat", "1") and props["hsvfilter"]["hue-shift"][4] == "mutable-playing"
    d = props["hsvdetector"]
    assert d["hue-var"][:4] == ("gfloat", "10", "0", "180") and d["saturation-var"][1:4] == ("0.150000006", "0", "1")
    assert d["value-var"][1:4] == ("0.300000012", "0", "1") and float(d["hue-ref"][2]) < -3e38
    assert props["roundedcorners"]["border-radius-px"] == ("guint", "0", "0", "4294967295", "mutable-playing")
    assert props["videocompare"]["hash-algo"][:2] == ("enum", "blockhash") and props["videocompare"]["max-dist-threshold"][:3] == ("gdouble", "0", "0")


def test_property_set_get_and_range_checks():
    el = gst.Element("hsvdetector")
    assert el.set_property("hue-var", 30) == 0 and el.get_property("hue-var") == "30"
    assert el.set_property("hue-var", 181) != 0 and el.get_property("hue-var") == "30"   # out of range: not set
    assert el.set_property("saturation-ref", -0.1) != 0 and el.set_property("nope", 1) != 0
    assert el.set_property("hue-ref", -1e30) == 0
    vc = gst.Element("videocompare")
    assert vc.set_property("hash-algo", "mean") == 0 and vc.get_property("hash-algo") == "mean"
    assert vc.set_property("hash-algo", 4) == 0 and vc.get_property("hash-algo") == "blockhash"
    assert vc.set_property("hash-algo", "dssim") != 0      # feature-gated in the reference, absent here
    assert vc.set_property("max-dist-threshold", -1) != 0
    rc = gst.Element("roundedcorners")
    assert rc.set_property("border-radius-px", 2.5) != 0 and rc.set_property("border-radius-px", 64) == 0
    cl = gst.Element("colorlut")
    assert cl.get_property("location") == "NULL"
    assert cl.set_property("location", "/tmp/x.cube") == 0 and cl.get_property("location") == "/tmp/x.cube"


def test_pad_templates_and_transform_caps():
    assert gst.Element("colorlut").pad_template_formats(gst.PAD_SINK) == ["RGBA64_LE", "RGBA64_BE", "RGBA"]
    hf = gst.Element("hsvfilter")
    assert hf.pad_template_formats(gst.PAD_SRC) == ["RGBx", "xRGB", "BGRx", "xBGR", "RGBA", "ARGB", "BGRA", "ABGR", "RGB", "BGR"]
    assert hf.transform_caps(gst.PAD_SINK, ["BGR", "I420"]) == ["BGR"]
    hd = gst.Element("hsvdetector")
    assert hd.pad_template_formats(gst.PAD_SINK) == ["RGBx", "xRGB", "BGRx", "xBGR", "RGB", "BGR"]
    assert hd.transform_caps(gst.PAD_SINK, ["BGRx"]) == ["RGBA", "ARGB", "BGRA", "ABGR"]
    assert hd.transform_caps(gst.PAD_SRC, ["RGBA"]) == ["RGBx", "xRGB", "BGRx", "xBGR", "RGB", "BGR"]
    rc = gst.Element("roundedcorners")
    assert rc.pad_template_formats(gst.PAD_SINK) == ["I420"] and rc.pad_template_formats(gst.PAD_SRC) == ["I420", "A420"]
    assert rc.transform_caps(gst.PAD_SINK, ["I420"]) == ["I420", "A420"]     # radius 0: both
    rc.set_property("border-radius-px", 10)
    assert rc.transform_caps(gst.PAD_SINK, ["I420"]) == ["A420"]             # radius != 0 forces A420
    assert rc.transform_caps(gst.PAD_SRC, ["A420"]) == ["I420"]
    assert gst.Element("videocompare").pad_template_formats(gst.PAD_SINK) == ["RGB", "RGBA"]


def test_videocompare_reference_pad_selection():
    vc = gst.Element("videocompare")
    assert vc.reference_pad == -1
    a, b, c = vc.request_pad(), vc.request_pad(), vc.request_pad()
    assert vc.reference_pad == a
    vc.release_pad(b)
    assert vc.reference_pad == a
    vc.release_pad(a)
    assert vc.reference_pad == c


def test_colorlut_start_errors_without_gpu_or_location():
    el = gst.Element("colorlut")
    assert el.start() != 0 and "location is not configured" in el.last_error   # imp.rs:175-180


# ---- pixel paths ------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_colorlut_element_pipeline(tmp_path):
    p = tmp_path / "mix.cube"
    p.write_text(synth.cube_text_3d(17, "mix", title="t"))
    el = gst.Element("colorlut")
    el.set_property("location", str(tmp_path / "missing.cube"))
    assert el.start() != 0 and "ResourceError::Read" in el.last_error           # imp.rs:182-187
    el.set_property("location", str(p))
    assert el.start() == 0
    cube = orc.cube_parse(p.read_text())
    for fmt, bpp in (("RGBA", 4), ("RGBA64_LE", 8), ("RGBA64_BE", 8)):
        w, h = 321, 33
        src = synth.frame_noise(fmt, w, h, 5, stride=bpp * w + 8)
        dst = np.full((h, bpp * w + 16), 0x5A, np.uint8)
        fin = gst.frame(fmt, w, h, [src], [src.shape[1]])
        fout = gst.frame(fmt, w, h, [dst], [dst.shape[1]])
        assert el.transform_frame(fin, fout) == gst.FLOW_OK
        assert (dst == orc.colorlut_apply(cube, fmt, w, h, src, dst_stride=dst.shape[1])).all()
    bad = gst.frame("BGRA", 4, 4, [np.zeros((4, 16), np.uint8)], [16])
    assert el.transform_frame(bad, bad) == gst.FLOW_NOT_NEGOTIATED
    assert el.stop() == 0
    f = gst.frame("RGBA", 4, 4, [np.zeros((4, 16), np.uint8)], [16])
    assert el.transform_frame(f, f) == gst.FLOW_ERROR and "No LUT configured" in el.last_error   # imp.rs:210-213


@pytest.mark.gpu
def test_hsvfilter_config1_gst_launch_equivalent():
    """gst-launch-1.0 videotestsrc ! video/x-raw,format=RGBA,width=640,height=480 ! hsvfilter hue-shift=90 ! ...
    (BASELINE config 1; the reference CPU path is the oracle)."""
    el = gst.Element("hsvfilter")
    el.set_property("hue-shift", 90)
    assert el.start() == 0
    w, h = 640, 480
    for frame in (synth.frame_ramps("RGBA", w, h), synth.frame_noise("RGBA", w, h, 0x5EED0001)):
        buf = frame.copy()
        assert el.transform_frame_ip(gst.frame("RGBA", w, h, [buf], [4 * w])) == gst.FLOW_OK
        assert (buf == orc.hsvfilter("RGBA", w, h, frame, hue_shift=90.0)).all()
    # properties are mutable in PLAYING: the next frame sees the new snapshot
    el.set_property("saturation-mul", 0.5)
    buf = frame.copy()
    el.transform_frame_ip(gst.frame("RGBA", w, h, [buf], [4 * w]))
    assert (buf == orc.hsvfilter("RGBA", w, h, frame, hue_shift=90.0, sat_mul=0.5)).all()
    el.stop()


@pytest.mark.gpu
def test_hsvdetector_then_roundedcorners_config3():
    """BASELINE config 3.  The two elements cannot be linked directly (hsvdetector emits RGBA, roundedcorners
    takes I420 -- SURVEY D1), so each is driven on its own 1920x1080 stream."""
    w, h = 1920, 1080
    det = gst.Element("hsvdetector")
    for k, v in (("hue-ref", 120), ("hue-var", 30), ("saturation-ref", 0.8), ("saturation-var", 0.2), ("value-ref", 0.8), ("value-var", 0.2)):
        assert det.set_property(k, v) == 0
    assert det.start() == 0
    src = synth.frame_noise("BGRx", w, h, 0x5EED0003)
    dst = np.zeros((h, 4 * w), np.uint8)
    assert det.transform_frame(gst.frame("BGRx", w, h, [src], [4 * w]), gst.frame("RGBA", w, h, [dst], [4 * w])) == gst.FLOW_OK
    assert (dst == orc.hsvdetector("BGRx", "RGBA", w, h, src, hue_ref=120.0, hue_var=30.0, sat_ref=0.8, sat_var=0.2,
                                   val_ref=0.8, val_var=0.2, threads=8)).all()
    det.stop()

    rc = gst.Element("roundedcorners")
    rc.set_property("border-radius-px", 64)
    assert rc.start() == 0
    assert rc.set_caps("I420", "A420", w, h) == 0 and not rc.passthrough
    y = np.zeros((h, w), np.uint8); u = np.zeros((h // 2, w // 2), np.uint8); v = u.copy()
    fin = gst.frame("I420", w, h, [y, u, v], [w, w // 2, w // 2])
    rcode, out = rc.prepare_output(fin)
    assert rcode == gst.FLOW_OK and out.n_planes == 4 and out.format == gst.FMT["A420"] and out.stride[3] == w
    mask = np.ctypeslib.as_array((np.ctypeslib.ctypes.c_uint8 * (w * h)).from_address(out.data[3])).reshape(h, w)
    assert (mask == orc.roundmask(w, h, w, 64)[:h]).all()
    assert out.data[0] == fin.data[0]                          # zero copy: same Y/U/V memories
    rcode2, out2 = rc.prepare_output(fin)
    assert out2.data[3] == out.data[3]                         # mask is shared, not regenerated per frame
    rc.set_property("border-radius-px", 0)                     # radius change -> regenerated: opaque plane
    rcode3, out3 = rc.prepare_output(fin)
    mask = np.ctypeslib.as_array((np.ctypeslib.ctypes.c_uint8 * (w * h)).from_address(out3.data[3]))
    assert (mask == 255).all()
    assert rc.set_caps("I420", "I420", w, h) == 0 and rc.passthrough
    rc.stop()


@pytest.mark.gpu
def test_videocompare_red_vs_red_posts_message_and_snow_does_not():
    """video/videofx/tests/videocompare.rs:57-139"""
    w, h = 320, 240
    vc = gst.Element("videocompare")
    assert vc.start() == 0
    ref, other = vc.request_pad(), vc.request_pad()
    red = synth.frame_solid("RGBA", w, h)
    red2 = red.copy()
    snow = synth.frame_noise("RGBA", w, h, 7)
    out = np.zeros_like(red)
    fr = lambda a: gst.frame("RGBA", w, h, [a], [4 * w])
    assert vc.aggregate_frames([fr(red), fr(red2)], [ref, other], 0, fr(out)) == gst.FLOW_OK
    msg = vc.pop_message()
    assert msg is not None and msg.startswith("videocompare, pad-distances=") and "pad\\=sink_1" in msg and "distance\\=(double)0" in msg
    assert "running-time=(guint64)0" in msg and (out == red).all()
    assert vc.aggregate_frames([fr(red), fr(snow)], [ref, other], 40_000_000, fr(out)) == gst.FLOW_OK
    assert vc.pop_message() is None                                        # distance > 0 with threshold 0: no message
    vc.set_property("max-dist-threshold", 64)
    assert vc.aggregate_frames([fr(red), fr(snow)], [ref, other], 80_000_000, None) == gst.FLOW_OK
    assert vc.pop_message() is not None
    small = synth.frame_solid("RGBA", 160, 120)
    assert vc.aggregate_frames([fr(red), gst.frame("RGBA", 160, 120, [small], [640])], [ref, other]) == gst.FLOW_NOT_NEGOTIATED
    assert vc.aggregate_frames([fr(red2)], [other]) == gst.FLOW_OK and vc.pop_message() is None   # reference pad has no buffer
    # every value of hash-algo (GstVideoCompareHashAlgorithm, videocompare/mod.rs:57-92) hashes: identical frames -> distance 0
    # -> message; snow vs red -> distance > 0 -> none at threshold 0.  Also on a size that is not a multiple of 8.
    vc.set_property("max-dist-threshold", 0)
    for algo in ("mean", "gradient", "vertgradient", "doublegradient", "blockhash"):
        vc.set_property("hash-algo", algo)
        assert vc.aggregate_frames([fr(red), fr(red2)], [ref, other]) == gst.FLOW_OK, algo
        m = vc.pop_message()
        assert m is not None and "distance\\=(double)0" in m, algo
        assert vc.aggregate_frames([fr(red), fr(snow)], [ref, other]) == gst.FLOW_OK
        assert vc.pop_message() is None, algo
        w2, h2 = 333, 241
        r2, s2 = synth.frame_solid("RGBA", w2, h2), synth.frame_noise("RGBA", w2, h2, 9)
        f2 = lambda a: gst.frame("RGBA", w2, h2, [a], [4 * w2])
        assert vc.aggregate_frames([f2(r2), f2(r2.copy())], [ref, other]) == gst.FLOW_OK and vc.pop_message() is not None, algo
        assert vc.aggregate_frames([f2(r2), f2(s2)], [ref, other]) == gst.FLOW_OK and vc.pop_message() is None, algo
    vc.stop()


@pytest.mark.gpu
def test_videocompare_with_device_resident_frames():
    """frames that live in HBM (memory:CUDAMemory-style pipeline): the element must not touch them with host code"""
    torch = pytest.importorskip("torch")
    w, h = 640, 480
    vc = gst.Element("videocompare")
    assert vc.start() == 0
    ref, other = vc.request_pad(), vc.request_pad()
    red = synth.frame_solid("RGBA", w, h)
    d_red, d_red2 = torch.from_numpy(red).cuda(), torch.from_numpy(red.copy()).cuda()
    d_snow = torch.from_numpy(synth.frame_noise("RGBA", w, h, 7)).cuda()
    d_out = torch.zeros_like(d_red)
    fr = lambda t: gst.frame("RGBA", w, h, [t], [4 * w])
    assert vc.aggregate_frames([fr(d_red), fr(d_red2)], [ref, other], 0, fr(d_out)) == gst.FLOW_OK
    msg = vc.pop_message()
    assert msg is not None and "distance\\=(double)0" in msg
    assert (d_out.cpu().numpy() == red).all()                     # output = the reference buffer, copied device to device
    host_out = np.zeros_like(red)
    assert vc.aggregate_frames([fr(d_red), fr(d_snow)], [ref, other], 1, fr(host_out)) == gst.FLOW_OK
    assert vc.pop_message() is None and (host_out == red).all()   # device reference -> host output buffer
    vc.stop()
