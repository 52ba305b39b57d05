"""GPU tests added in round 2: the round-1 review findings (1D LUT geometry, mixed host/device ordering) and the new kernels."""
import numpy as np
import pytest

import b200vfx
import oracle_binding as orc
from b200vfx import synth

pytestmark = pytest.mark.gpu
NT = 8


@pytest.fixture(scope="module")
def ctx():
    c = b200vfx.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("w,h,pad", [(1920, 1080, 0), (3840, 2160, 0), (1918, 1080, 0), (1919, 1081, 12), (2562, 1440, 0)])
@pytest.mark.parametrize("stream_path", [0, -1])
def test_colorlut_1d_device_frames_any_size(ctx, w, h, pad, stream_path):
    """1D LUT on device frames beyond 592*256*8 pixels: every pixel must be written (the non-TMA kernel is item-persistent),
    with stream_path = 0 and with widths that are not a multiple of 4 (rows then are not flattened / not 16-byte aligned)"""
    torch = pytest.importorskip("torch")
    cube = orc.cube_parse(synth.cube_text_1d(1024, 2.0))
    ctx.colorlut_set_lut(cube.kind, cube.size, cube.values, cube.scale, cube.offset)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    ctx.set_option("stream_path", stream_path)
    try:
        stride = 4 * w + pad
        frame = synth.frame_noise("RGBA", w, h, 0x1D + w, stride=stride)
        exp = orc.colorlut_apply(cube, "RGBA", w, h, frame, threads=NT)
        d_src = torch.from_numpy(frame).cuda()
        for px in (4, 8, 16):
            ctx.set_option("memo_px", px)
            d_dst = torch.full_like(d_src, 0xA5)
            ctx.colorlut_process("RGBA", w, h, d_src, stride, d_dst, stride)
            torch.cuda.synchronize()
            got = d_dst.cpu().numpy()
            assert (got[:, :4 * w] == exp[:, :4 * w]).all(), (w, h, px)
            assert (got[:, 4 * w:] == 0xA5).all()          # padding untouched
    finally:
        ctx.set_option("stream_path", -1)
        ctx.set_option("memo_px", 8)


def test_mixed_host_device_calls_follow_the_context_stream():
    """a device frame produced asynchronously on the context stream and consumed by a call whose OTHER plane is host memory
    (which runs on the library's internal stream) must be ordered after its producer"""
    torch = pytest.importorskip("torch")
    w, h = 3840, 2160
    frame = synth.frame_noise("BGRx", w, h, 321)
    fo = dict(hue_shift=33.0)
    step1 = orc.hsvfilter("BGRx", w, h, frame.copy(), threads=NT, **fo)
    dkw = dict(hue_ref=120.0, hue_var=60.0, saturation_ref=0.6, saturation_var=0.4, value_ref=0.6, value_var=0.4)
    exp = orc.hsvdetector("BGRx", "RGBA", w, h, step1, hue_ref=120.0, hue_var=60.0, sat_ref=0.6, sat_var=0.4, val_ref=0.6, val_var=0.4, threads=NT)
    exp_sums = orc.blockhash_sums("RGBA", w, h, exp)
    with b200vfx.Context(0) as ctx:
        s = torch.cuda.Stream()
        ctx.set_stream(s.cuda_stream)
        ctx.set_option("hsv_memo", 0)      # the slow direct kernel: the producer is certainly still running when the consumer is enqueued
        for rep in range(3):
            with torch.cuda.stream(s):
                d = torch.from_numpy(frame).cuda(non_blocking=False)
            for _ in range(2):             # queue some work in front of the producer
                ctx.hsvfilter_process("BGRx", w, h, torch.from_numpy(frame).cuda(), 4 * w, hue_shift=1.0)
            ctx.hsvfilter_process("BGRx", w, h, d, 4 * w, **fo)                                  # device, in place, async on s
            out = np.zeros((h, 4 * w), np.uint8)
            ctx.hsvdetector_process("BGRx", "RGBA", w, h, d, 4 * w, out, 4 * w, **dkw)          # device src, HOST dst
            assert (out == exp).all(), rep
            # blockhash batch mixing a device frame that is still being produced with a host frame
            with torch.cuda.stream(s):
                dd = torch.zeros((h, 4 * w), dtype=torch.uint8, device="cuda")
            ctx.hsvdetector_process("BGRx", "RGBA", w, h, d, 4 * w, dd, 4 * w, **dkw)            # device -> device, async on s
            sums = np.zeros(128, np.uint32)
            ctx.blockhash_sums_batch("RGBA", w, h, [dd, exp], [4 * w, 4 * w], sums)
            assert (sums[:64] == exp_sums).all() and (sums[64:] == exp_sums).all(), rep
            s.synchronize()


# ---- direct (non-memoised) hsv kernels: the branch-free arithmetic of hsv_fast.cuh on the GPU ------------------------
from test_hsv_fast_model import FILTER_SETTINGS, DETECT_SETTINGS, all_colors_frame   # noqa: E402


@pytest.mark.parametrize("st", FILTER_SETTINGS)
def test_hsvfilter_direct_kernel_all_colors_all_setting_classes(ctx, st):
    """every 24-bit colour through the direct kernel (hsv_memo = 0) == oracle, for ordinary settings and for the ones that
    select the general code (NaN / inf / huge / near-denormal hue-shift)"""
    torch = pytest.importorskip("torch")
    ctx.set_option("hsv_memo", 0)
    try:
        frame = all_colors_frame()
        kw = dict(hue_shift=st[0], sat_mul=st[1], sat_off=st[2], val_mul=st[3], val_off=st[4])
        exp = orc.hsvfilter("RGBx", 4096, 4096, frame.copy(), threads=NT, **kw)
        ctx.set_stream(torch.cuda.current_stream().cuda_stream)
        d = torch.from_numpy(frame).cuda()
        ctx.hsvfilter_process("RGBx", 4096, 4096, d, 4 * 4096, hue_shift=st[0], saturation_mul=st[1], saturation_off=st[2],
                              value_mul=st[3], value_off=st[4])
        torch.cuda.synchronize()
        assert (d.cpu().numpy() == exp).all(), st
    finally:
        ctx.set_option("hsv_memo", -1)


@pytest.mark.parametrize("st", DETECT_SETTINGS)
@pytest.mark.parametrize("ifmt,ofmt", [("RGBx", "RGBA"), ("xBGR", "ARGB")])
def test_hsvdetector_direct_kernel_all_colors(ctx, st, ifmt, ofmt):
    torch = pytest.importorskip("torch")
    ctx.set_option("hsv_memo", 0)
    try:
        frame = all_colors_frame()
        if ifmt == "xBGR":
            frame = np.ascontiguousarray(frame.reshape(-1, 4)[:, ::-1]).reshape(4096, -1)   # x,B,G,R <- R,G,B,x reversed
        exp = orc.hsvdetector(ifmt, ofmt, 4096, 4096, frame, hue_ref=st[0], hue_var=st[1], sat_ref=st[2], sat_var=st[3],
                              val_ref=st[4], val_var=st[5], threads=NT)
        ctx.set_stream(torch.cuda.current_stream().cuda_stream)
        d = torch.from_numpy(frame).cuda()
        o = torch.zeros_like(d)
        ctx.hsvdetector_process(ifmt, ofmt, 4096, 4096, d, 4 * 4096, o, 4 * 4096, hue_ref=st[0], hue_var=st[1],
                                saturation_ref=st[2], saturation_var=st[3], value_ref=st[4], value_var=st[5])
        torch.cuda.synchronize()
        assert (o.cpu().numpy() == exp).all(), st
    finally:
        ctx.set_option("hsv_memo", -1)


@pytest.mark.parametrize("fmt", ["RGB", "BGR", "ARGB", "BGRx"])
def test_hsvfilter_direct_kernel_other_formats_and_strides(ctx, fmt):
    """3-byte pixels and unaligned rows take the one-pixel-per-thread kernel with the same arithmetic"""
    ctx.set_option("hsv_memo", 0)
    try:
        bpp = 3 if fmt in ("RGB", "BGR") else 4
        for (w, h, pad, off) in ((641, 37, 5, 0), (1280, 64, 0, 0), (333, 17, 3, 1)):
            stride = ((w * bpp + 3) // 4) * 4 + pad
            base = synth.frame_noise(fmt, w, h, 0xF00 + w, stride=stride)
            buf = np.zeros(stride * h + 8, np.uint8)
            view = buf[off:off + stride * h].reshape(h, stride)
            view[:] = base
            exp = orc.hsvfilter(fmt, w, h, base.copy(), hue_shift=200.0, sat_mul=1.2, val_off=-0.05, threads=NT)
            ctx.hsvfilter_process(fmt, w, h, view, stride, hue_shift=200.0, saturation_mul=1.2, value_off=-0.05)
            assert (view[:, :w * bpp] == exp[:, :w * bpp]).all(), (fmt, w, h)
            assert (view[:, w * bpp:] == base[:, w * bpp:]).all()
    finally:
        ctx.set_option("hsv_memo", -1)


def test_pdl_small_frames_ring_of_three_outputs_stress(ctx):
    """small frames (non-lingering kernels: nothing bounds how many stay co-resident) with a ring of only 3 output buffers:
    20 000 back-to-back launches with per-frame distinct content must leave exactly the right frames in the ring -- the
    admission rule turns every reuse of a possibly-live buffer into a plain (fully ordered) launch"""
    torch = pytest.importorskip("torch")
    w, h = 640, 480
    cube = orc.cube_parse(synth.cube_text_3d(17, "mix"))
    ctx.colorlut_set_lut(cube.kind, cube.size, cube.values, cube.scale, cube.offset)
    ctx.colorlut_set_mode(0)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    nin = 7
    frames = [synth.frame_noise("RGBA", w, h, 1000 + i) for i in range(nin)]
    exp = [orc.colorlut_apply(cube, "RGBA", w, h, f, threads=NT) for f in frames]
    d_in = [torch.from_numpy(f).cuda() for f in frames]
    d_out = [torch.zeros((h, 4 * w), dtype=torch.uint8, device="cuda") for _ in range(3)]
    n = 20000
    for i in range(n):
        ctx.colorlut_process("RGBA", w, h, d_in[i % nin], 4 * w, d_out[i % 3], 4 * w)
        if i % 4999 == 4998:                      # spot checks in flight: the last three launches' outputs
            torch.cuda.synchronize()
            for k in range(3):
                j = i - k
                assert (d_out[j % 3].cpu().numpy() == exp[j % nin]).all(), (i, k)
    torch.cuda.synchronize()
    for k in range(3):
        j = n - 1 - k
        assert (d_out[j % 3].cpu().numpy() == exp[j % nin]).all(), k
    # the same with in-place hsvfilter (memoised) chained on the ring: y = f(x) then filter(y) in place, 3 buffers
    ctx.set_option("hsv_memo", 1)
    try:
        flt = [orc.hsvfilter("RGBA", w, h, e.copy(), hue_shift=33.0, threads=NT) for e in exp]
        for i in range(3000):
            ctx.colorlut_process("RGBA", w, h, d_in[i % nin], 4 * w, d_out[i % 3], 4 * w)
            ctx.hsvfilter_process("RGBA", w, h, d_out[i % 3], 4 * w, hue_shift=33.0)
        torch.cuda.synchronize()
        for k in range(3):
            j = 2999 - k
            assert (d_out[j % 3].cpu().numpy() == flt[j % nin]).all(), k
    finally:
        ctx.set_option("hsv_memo", -1)


@pytest.mark.parametrize("fmt", ["RGBA64_LE", "RGBA64_BE"])
@pytest.mark.parametrize("lut", ["mix33", "domain17", "ident2", "mix65"])
def test_colorlut_rgba64_four_pixels_per_thread_kernel(ctx, fmt, lut):
    """RGBA64 + 3D LUT through colorlut_direct64x4_kernel (LUT cell cached in registers across 4 consecutive pixels, packed
    exact products) == oracle == the one-pixel-per-thread kernel, on coherent, noisy and boundary-heavy content"""
    torch = pytest.importorskip("torch")
    text = {"mix33": synth.cube_text_3d(33, "mix"), "domain17": synth.cube_text_3d(17, "mix", domain=((-0.25, -0.1, 0.0), (1.5, 1.2, 1.0))),
            "ident2": synth.cube_text_3d(2, "identity"), "mix65": synth.cube_text_3d(65, "mix")}[lut]
    cube = orc.cube_parse(text)
    ctx.colorlut_set_lut(cube.kind, cube.size, cube.values, cube.scale, cube.offset)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    for (w, h, kind) in ((1920, 64, "ramps"), (1024, 33, "noise"), (640, 17, "natural"), (256, 256, "edges"), (1022, 5, "noise")):
        if kind == "ramps":
            frame = synth.frame_ramps(fmt, w, h)
        elif kind == "natural":   # 16-bit ramps with +-200 of noise: neighbouring pixels mostly share a LUT cell, not always
            dt = "<u2" if fmt.endswith("LE") else ">u2"
            base = synth.frame_ramps(fmt, w, h).view(dt).astype(np.int64)
            noise = np.random.default_rng(9).integers(-200, 201, base.shape)
            frame = np.clip(base + noise, 0, 65535).astype(dt).view(np.uint8).reshape(h, 8 * w)
        elif kind == "edges":   # every 16-bit channel value that sits on or next to a LUT cell boundary, plus extremes
            v = np.array(sorted(set([0, 1, 2, 65533, 65534, 65535] + [int(round(k * 65535 / 32)) + d for k in range(33) for d in (-1, 0, 1)])), np.int64)
            v = v[(v >= 0) & (v <= 65535)].astype(np.uint16)
            rng = np.random.default_rng(1)
            px = np.stack([rng.choice(v, w * h), rng.choice(v, w * h), rng.choice(v, w * h), rng.integers(0, 65536, w * h).astype(np.uint16)], axis=1)
            frame = px.astype("<u2" if fmt.endswith("LE") else ">u2").view(np.uint8).reshape(h, 8 * w)
        else:
            frame = synth.frame_noise(fmt, w, h, 77 + w)
        if frame is None:
            continue
        exp = orc.colorlut_apply(cube, fmt, w, h, frame, threads=NT)
        d = torch.from_numpy(np.ascontiguousarray(frame)).cuda()
        for x4 in (1, 0):
            ctx.set_option("rgba64_x4", x4)
            o = torch.zeros_like(d)
            ctx.colorlut_process(fmt, w, h, d, 8 * w, o, 8 * w)
            torch.cuda.synchronize()
            assert (o.cpu().numpy() == exp).all(), (fmt, lut, w, h, kind, x4)
        ctx.set_option("rgba64_x4", 1)
        out = np.zeros_like(frame)
        ctx.colorlut_process(fmt, w, h, frame, 8 * w, out, 8 * w)      # host frames (staged)
        assert (out == exp).all(), (fmt, lut, w, h, kind, "host")


# ---- asynchronous host-frame mode (b200vfx_ctx_set_host_async + fences) --------------------------------------------------
@pytest.mark.parametrize("zero_copy", [0, 1, 2])
def test_host_async_colorlut_stream_matches_sync(zero_copy):
    """frames submitted from pinned host memory without waiting for each other: every output equals the synchronous call's,
    with the staged copy-engine pipeline (two staging slots), the zero-copy kernel, and the auto-probe in between"""
    torch = pytest.importorskip("torch")
    w, h = 1920, 1080
    cube = orc.cube_parse(synth.cube_text_3d(17, "mix"))
    frames = [synth.frame_noise("RGBA", w, h, 500 + i) if i % 2 else synth.frame_natural("RGBA", w, h, 500 + i) for i in range(5)]
    exp = [orc.colorlut_apply(cube, "RGBA", w, h, f, threads=8) for f in frames]
    with b200vfx.Context(0) as ctx:
        ctx.colorlut_set_lut(cube.kind, cube.size, cube.values, cube.scale, cube.offset)
        ctx.set_option("zero_copy", zero_copy)
        ring = 4
        h_in = [torch.empty((h, 4 * w), dtype=torch.uint8).pin_memory() for _ in range(ring)]
        h_out = [torch.empty((h, 4 * w), dtype=torch.uint8).pin_memory() for _ in range(ring)]
        ctx.set_host_async(True)
        fences, checked = [], 0
        n = 23
        for i in range(n):
            if i >= 3:                      # at most three frames in flight: frame i - 3 is done, its buffers are ours again
                fences[i - 3].wait()
                assert (h_out[(i - 3) % ring].numpy() == exp[(i - 3) % 5]).all(), i
                checked += 1
            h_in[i % ring].numpy()[:] = frames[i % 5]
            h_out[i % ring].numpy()[:] = 0x5A
            ctx.colorlut_process("RGBA", w, h, h_in[i % ring].numpy(), 4 * w, h_out[i % ring].numpy(), 4 * w)
            fences.append(ctx.fence())
        ctx.synchronize()
        assert all(f.done() for f in fences)
        for i in range(n - 3, n):
            assert (h_out[i % ring].numpy() == exp[i % 5]).all(), i
        assert checked == n - 3
        ctx.set_host_async(False)           # back to synchronous calls on the same context
        out = np.zeros_like(frames[0])
        ctx.colorlut_process("RGBA", w, h, frames[2], 4 * w, out, 4 * w)
        assert (out == exp[2]).all()
        for f in fences:
            f.close()


def test_host_async_in_place_and_two_plane_elements():
    """hsvfilter (in place), hsvdetector and convert_packed in asynchronous mode, strides with padding, frame sizes changing
    between calls (staging slots regrow)"""
    torch = pytest.importorskip("torch")
    with b200vfx.Context(0) as ctx:
        ctx.set_host_async(True)
        ctx.set_option("hsv_memo", 0)
        jobs = []
        for k, (w, h, pad) in enumerate(((640, 360, 0), (1280, 720, 16), (322, 100, 0), (1280, 720, 16), (640, 360, 0), (1920, 1080, 0))):
            f = synth.frame_noise("RGBA", w, h, 40 + k, stride=4 * w + pad)
            a = torch.from_numpy(f.copy()).pin_memory()
            ctx.hsvfilter_process("RGBA", w, h, a.numpy(), 4 * w + pad, hue_shift=15.0 * k, saturation_mul=1.1)
            b_in = torch.from_numpy(f.copy()).pin_memory()
            b_out = torch.zeros((h, 4 * w + pad), dtype=torch.uint8).pin_memory()
            ctx.hsvdetector_process("BGRx", "BGRA", w, h, b_in.numpy(), 4 * w + pad, b_out.numpy(), 4 * w + pad, hue_ref=40.0 * k, hue_var=60.0,
                                    saturation_ref=0.5, saturation_var=0.5, value_ref=0.5, value_var=0.5)
            c_out = torch.zeros((h, 4 * w), dtype=torch.uint8).pin_memory()
            ctx.convert_packed("RGBA", "xBGR", w, h, b_in.numpy(), 4 * w + pad, c_out.numpy(), 4 * w)
            jobs.append((k, w, h, pad, f, a, b_in, b_out, c_out, ctx.fence()))
        import np_convert as npc
        for k, w, h, pad, f, a, b_in, b_out, c_out, fence in jobs:
            fence.wait()
            assert (a.numpy()[:, :4 * w] == orc.hsvfilter("RGBA", w, h, f, hue_shift=15.0 * k, sat_mul=1.1)[:, :4 * w]).all(), k
            expd = orc.hsvdetector("BGRx", "BGRA", w, h, f, hue_ref=40.0 * k, hue_var=60.0, sat_ref=0.5, sat_var=0.5, val_ref=0.5, val_var=0.5)
            assert (b_out.numpy()[:, :4 * w] == expd[:, :4 * w]).all(), k
            assert (c_out.numpy() == npc.convert_packed("RGBA", "xBGR", w, h, f)).all(), k
            fence.close()
        ctx.synchronize()


def test_host_async_edge_cases():
    """a fence with nothing submitted is reached at once; pageable frames work in asynchronous mode (the driver stages them, so
    the call is effectively synchronous); device-pointer calls are untouched by the mode; toggling the mode drains"""
    torch = pytest.importorskip("torch")
    w, h = 320, 96
    cube = orc.cube_parse(synth.cube_text_1d(32))
    frame = synth.frame_noise("RGBA", w, h, 3)
    exp = orc.colorlut_apply(cube, "RGBA", w, h, frame)
    with b200vfx.Context(0) as ctx:
        f0 = ctx.fence(); f0.wait(); assert f0.done(); f0.close()          # nothing submitted, mode off
        ctx.colorlut_set_lut(cube.kind, cube.size, cube.values, cube.scale, cube.offset)
        ctx.set_host_async(True)
        out = np.zeros_like(frame)
        ctx.colorlut_process("RGBA", w, h, frame, 4 * w, out, 4 * w)       # pageable numpy arrays
        f1 = ctx.fence(); f1.wait(); f1.close()
        assert (out == exp).all()
        d_in, d_out = torch.from_numpy(frame).cuda(), torch.zeros((h, 4 * w), dtype=torch.uint8, device="cuda")
        ctx.set_stream(torch.cuda.current_stream().cuda_stream)
        ctx.colorlut_process("RGBA", w, h, d_in, 4 * w, d_out, 4 * w)      # device frames: the context stream, as always
        torch.cuda.synchronize()
        assert (d_out.cpu().numpy() == exp).all()
        pin_in, pin_out = torch.from_numpy(frame).pin_memory(), torch.zeros((h, 4 * w), dtype=torch.uint8).pin_memory()
        for _ in range(5):
            ctx.colorlut_process("RGBA", w, h, pin_in.numpy(), 4 * w, pin_out.numpy(), 4 * w)
        ctx.set_host_async(False)                                            # drains what is in flight
        assert (pin_out.numpy() == exp).all()
        sums = np.zeros(64, np.uint32)
        ctx.set_host_async(True)
        ctx.colorlut_process("RGBA", w, h, pin_in.numpy(), 4 * w, pin_out.numpy(), 4 * w)
        ctx.blockhash_sums("RGBA", w, h, frame, 4 * w, sums)                # a reduction stays synchronous in the mode
        assert (sums == orc.blockhash_sums("RGBA", w, h, frame)).all()
        ctx.synchronize()
        assert (pin_out.numpy() == exp).all()


def test_host_async_mixed_planes_are_ordered_with_the_context_stream():
    """asynchronous mode, host frame in, DEVICE frame out: the kernel runs on the internal stream and nothing waits for it
    before the call returns -- a device-pointer call that follows (context stream) must still see its result"""
    torch = pytest.importorskip("torch")
    w, h = 1920, 1080
    cube = orc.cube_parse(synth.cube_text_3d(17, "mix"))
    frames = [synth.frame_noise("BGRx", w, h, 900 + i) for i in range(6)]
    kw = dict(hue_ref=200.0, hue_var=90.0, saturation_ref=0.5, saturation_var=0.5, value_ref=0.5, value_var=0.5)
    okw = dict(hue_ref=200.0, hue_var=90.0, sat_ref=0.5, sat_var=0.5, val_ref=0.5, val_var=0.5)
    with b200vfx.Context(0) as ctx:
        ctx.set_stream(torch.cuda.current_stream().cuda_stream)
        ctx.set_option("hsv_memo", 0)
        ctx.colorlut_set_lut(cube.kind, cube.size, cube.values, cube.scale, cube.offset)
        ctx.set_host_async(True)
        pins = [torch.from_numpy(f).pin_memory() for f in frames]
        mid = torch.zeros((h, 4 * w), dtype=torch.uint8, device="cuda")
        outs = [torch.zeros((h, 4 * w), dtype=torch.uint8, device="cuda") for _ in frames]
        for i, p in enumerate(pins):
            # host -> device on the internal stream (asynchronous), then device -> device on the context stream; `mid` is
            # rewritten by the next iteration's first call while this iteration's second call may still be reading it
            ctx.hsvdetector_process("BGRx", "RGBA", w, h, p.numpy(), 4 * w, mid, 4 * w, **kw)
            ctx.colorlut_process("RGBA", w, h, mid, 4 * w, outs[i], 4 * w)
        ctx.synchronize()
        torch.cuda.synchronize()
        for i, f in enumerate(frames):
            exp = orc.colorlut_apply(cube, "RGBA", w, h, orc.hsvdetector("BGRx", "RGBA", w, h, f, **okw))
            assert (outs[i].cpu().numpy() == exp).all(), i
