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


@pytest.mark.parametrize("variant", ["tma", "blocks", "rows", "rows_2_per_sm"])
@pytest.mark.parametrize("w,h,pad", [(3840, 2160, 0), (7680, 4320, 0), (1920, 1080, 64), (64, 48, 0), (32, 8, 16), (4096, 16, 0)])
def test_blockhash_kernel_variants(ctx, variant, w, h, pad):
    """the three decompositions of the block sums: TMA-fed tiles, one CTA per hash block, whole-row streaming CTAs"""
    torch = pytest.importorskip("torch")
    ctx.set_option("blockhash_tma", 1 if variant == "tma" else 0)
    ctx.set_option("blockhash_rows", {"rows": 2, "rows_2_per_sm": 1}.get(variant, 0))
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
        ctx.set_option("blockhash_tma", 0)
        ctx.set_option("blockhash_rows", 2)


@pytest.mark.parametrize("rows", [2, 1, 0])
@pytest.mark.parametrize("w,h,hw,hh,n", [(3840, 2160, 16, 16, 1), (1920, 1080, 4, 4, 2), (640, 480, 16, 8, 3), (256, 256, 64, 64, 1),
                                         (1280, 720, 80, 16, 2), (4096, 64, 256, 2, 1), (2048, 4, 512, 4, 1)])
def test_blockhash_other_grids(ctx, rows, w, h, hw, hh, n):
    """hash grids other than 8x8 (image_hasher's hash_size), batches, block widths down to one 16-byte vector; a grid wider
    than the row kernel's shared bins falls back to the per-block kernel"""
    torch = pytest.importorskip("torch")
    ctx.set_option("blockhash_tma", 0)
    ctx.set_option("blockhash_rows", rows)
    try:
        frames = [synth.frame_noise("RGBA", w, h, 900 + i) for i in range(n)]
        frames[0][::3, 3::20] = 0
        exp = np.concatenate([orc.blockhash_sums("RGBA", w, h, f, hw, hh) for f in frames])
        dev = [torch.from_numpy(f).cuda() for f in frames]
        ds = torch.zeros(hw * hh * n, dtype=torch.int32, device="cuda")
        ctx.set_stream(torch.cuda.current_stream().cuda_stream)
        for _ in range(2):
            ctx.blockhash_sums_batch("RGBA", w, h, dev, [4 * w] * n, ds, hw=hw, hh=hh)
        torch.cuda.synchronize()
        assert (ds.cpu().numpy().view(np.uint32) == exp).all()
    finally:
        ctx.set_option("blockhash_tma", 0)
        ctx.set_option("blockhash_rows", 2)


@pytest.mark.parametrize("rows", [2, 1, 0])
def test_blockhash_back_to_back_launches_overlap_safely(ctx, rows):
    """consecutive block-sum launches overlap (programmatic dependent launch: the frame is read while the previous launch
    drains, scratch and sums are touched after griddepcontrol.wait): results of a long unsynchronised train, and a frame
    produced by the launch just before (colorlut -> blockhash on its output) must still be exact"""
    torch = pytest.importorskip("torch")
    w, h = 1920, 1080
    ctx.set_option("blockhash_tma", 0)
    ctx.set_option("blockhash_rows", rows)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    try:
        frames = [synth.frame_noise("RGBA", w, h, 300 + i) for i in range(4)]
        exp = [orc.blockhash_sums("RGBA", w, h, f) for f in frames]
        dev = [torch.from_numpy(f).cuda() for f in frames]
        sums = [torch.zeros(64, dtype=torch.int32, device="cuda") for _ in range(5)]
        for i in range(60):
            ctx.blockhash_sums("RGBA", w, h, dev[i % 4], 4 * w, sums[i % 5])
        torch.cuda.synchronize()
        for k in range(5):
            last = max(i for i in range(60) if i % 5 == k)
            assert (sums[k].cpu().numpy().view(np.uint32) == exp[last % 4]).all(), k
        # producer -> consumer: the frame being hashed is written by the kernel launched just before
        cube = orc.cube_parse(synth.cube_text_3d(9, "mix"))
        ctx.colorlut_set_lut(cube.kind, cube.size, cube.values, cube.scale, cube.offset)
        graded = [orc.colorlut_apply(cube, "RGBA", w, h, f) for f in frames]
        exp2 = [orc.blockhash_sums("RGBA", w, h, g) for g in graded]
        out = torch.empty_like(dev[0])
        got = [torch.zeros(64, dtype=torch.int32, device="cuda") for _ in range(12)]
        for i in range(12):
            ctx.colorlut_process("RGBA", w, h, dev[i % 4], 4 * w, out, 4 * w)   # overwrites the frame the previous hash is reading
            ctx.blockhash_sums("RGBA", w, h, out, 4 * w, got[i])
        torch.cuda.synchronize()
        for i in range(12):
            assert (got[i].cpu().numpy().view(np.uint32) == exp2[i % 4]).all(), i
    finally:
        ctx.set_option("blockhash_rows", 2)


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


@pytest.mark.parametrize("content", ["natural", "ramps", "noise", "mixed"])
def test_memo_tile_kernel_matches_oracle(ctx, content):
    """memo_tile_kernel (option memo_tile=1): per-tile shared-memory copy of the colour sub-cube when it fits, direct
    gathers otherwise -- decided per 64x64 tile; partial tiles at the right/bottom edge, padded strides, host + device"""
    torch = pytest.importorskip("torch")
    cube = orc.cube_parse(synth.cube_text_3d(33, "mix"))
    ctx.colorlut_set_lut(cube.kind, cube.size, cube.values, cube.scale, cube.offset)
    ctx.set_option("memo_tile", 1)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    try:
        for (w, h, pad) in ((3840, 2160, 0), (1920, 1080, 0), (100, 70, 16), (64, 64, 0), (4, 1, 0), (1000, 3, 48), (260, 129, 0)):
            stride = 4 * w + pad
            if content == "natural":
                frame = synth.frame_natural("RGBA", w, h, 11, amp=3)
            elif content == "ramps":
                frame = synth.frame_ramps("RGBA", w, h)
            elif content == "noise":
                frame = synth.frame_noise("RGBA", w, h, 12)
            else:  # left half smooth (cached tiles), right half random (gather tiles), one constant tile
                frame = synth.frame_natural("RGBA", w, h, 13, amp=2)
                nz = synth.frame_noise("RGBA", w, h, 14)
                frame[:, 2 * w:] = nz[:, 2 * w:4 * w]
                frame[: min(h, 64), : min(4 * w, 256)] = 0x40
            src = np.full((h, stride), 0xA5, np.uint8)
            src[:, :4 * w] = frame[:, :4 * w]
            exp = orc.colorlut_apply(cube, "RGBA", w, h, src, dst_stride=stride, threads=8)
            d_in = torch.from_numpy(src).cuda()
            d_out = torch.full((h, stride), 0x5A, dtype=torch.uint8, device="cuda")
            ctx.colorlut_process("RGBA", w, h, d_in, stride, d_out, stride)
            torch.cuda.synchronize()
            assert (d_out.cpu().numpy() == exp).all(), (w, h, pad)       # includes the untouched padding (0x5A)
    finally:
        ctx.set_option("memo_tile", 0)


@pytest.mark.parametrize("fmt", ["RGBA", "xRGB", "BGRx", "ABGR"])
def test_memo_tile_kernel_hsvfilter_in_place(ctx, fmt):
    torch = pytest.importorskip("torch")
    kw = dict(hue_shift=33.0, saturation_mul=1.1, value_off=-0.05)
    okw = dict(hue_shift=33.0, sat_mul=1.1, val_off=-0.05)
    ctx.set_option("memo_tile", 1)
    ctx.set_option("hsv_memo", 1)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    try:
        for (w, h) in ((1920, 1080), (132, 70), (64, 5)):
            for frame in (synth.frame_natural(fmt, w, h, 5, amp=3), synth.frame_noise(fmt, w, h, 6)):
                exp = orc.hsvfilter(fmt, w, h, frame.copy(), threads=8, **okw)
                d = torch.from_numpy(frame).cuda()
                ctx.hsvfilter_process(fmt, w, h, d, frame.shape[1], **kw)
                torch.cuda.synchronize()
                assert (d.cpu().numpy() == exp).all(), (fmt, w, h)
    finally:
        ctx.set_option("memo_tile", 0)
        ctx.set_option("hsv_memo", -1)


@pytest.mark.parametrize("ctas", [2, 3, 4, 8])
def test_persistent_lookup_kernels_any_grid_cap(ctx, ctas):
    """memo_ctas caps the persistent grid of the table-lookup kernels (CTAs per SM): every cap gives the same bytes, for
    packed frames (flattened to one row of items), padded strides (row x chunk items) and frames smaller than one CTA"""
    torch = pytest.importorskip("torch")
    cube = orc.cube_parse(synth.cube_text_3d(17, "mix"))
    ctx.colorlut_set_lut(cube.kind, cube.size, cube.values, cube.scale, cube.offset)
    ctx.set_option("memo_ctas", ctas)
    ctx.set_option("hsv_memo", 1)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    try:
        for (w, h, pad) in ((3840, 2160, 0), (3840, 700, 64), (5000, 3, 0), (33, 1, 4), (2049, 257, 12)):
            stride = 4 * w + pad
            src = np.full((h, stride), 0xA5, np.uint8)
            src[:, :4 * w] = synth.frame_noise("RGBA", w, h, 31 + w)[:, :4 * w]
            exp = orc.colorlut_apply(cube, "RGBA", w, h, src, dst_stride=stride, threads=8)
            d_in = torch.from_numpy(src).cuda()
            outs = [torch.full((h, stride), 0x5A, dtype=torch.uint8, device="cuda") for _ in range(3)]
            for o in outs:                                   # back to back: consecutive frames overlap (PDL)
                ctx.colorlut_process("RGBA", w, h, d_in, stride, o, stride)
            torch.cuda.synchronize()
            for o in outs:
                assert (o.cpu().numpy() == exp).all(), (w, h, pad)
            expf = orc.hsvfilter("RGBA", w, h, src.copy(), hue_shift=45.0, threads=8)
            d = torch.from_numpy(src).cuda()
            ctx.hsvfilter_process("RGBA", w, h, d, stride, hue_shift=45.0)
            torch.cuda.synchronize()
            assert (d.cpu().numpy() == expf).all(), (w, h, pad)
    finally:
        ctx.set_option("memo_ctas", 4)
        ctx.set_option("hsv_memo", -1)


def test_device_resident_chain_one_upload_one_download():
    """hsvfilter -> hsvdetector -> colordetect + blockhash on frames that stay in HBM (library-owned
    device buffers, one upload at the head, one download at the tail, all elements on one stream) == the oracle chain."""
    w, h = 1920, 1080
    src = synth.frame_natural("BGRx", w, h, 77, amp=4)
    fkw = dict(hue_shift=40.0, saturation_mul=1.1)
    dkw = dict(hue_ref=120.0, hue_var=60.0, saturation_ref=0.6, saturation_var=0.4, value_ref=0.6, value_var=0.4)
    step1 = orc.hsvfilter("BGRx", w, h, src.copy(), hue_shift=40.0, sat_mul=1.1, threads=8)
    step2 = orc.hsvdetector("BGRx", "RGBA", w, h, step1, hue_ref=120.0, hue_var=60.0, sat_ref=0.6, sat_var=0.4, val_ref=0.6, val_var=0.4, threads=8)
    with b200vfx.Context(0) as ctx:
        ctx.set_option("hsv_memo", 1)
        nb = 4 * w * h
        d_a, d_b = ctx.device_alloc(nb), ctx.device_alloc(nb)
        d_hist, d_sums = ctx.device_alloc(4 * 32768), ctx.device_alloc(4 * 64)
        host_in = np.ascontiguousarray(src)
        out = np.zeros((h, 4 * w), np.uint8)
        hist, sums = np.zeros(32768, np.uint32), np.zeros(64, np.uint32)
        for _ in range(3):   # repeated: frames of a stream reuse the same device buffers
            ctx.upload(d_a, 4 * w, host_in, 4 * w, 4 * w, h)
            ctx.hsvfilter_process("BGRx", w, h, d_a, 4 * w, **fkw)                               # in place, device
            ctx.hsvdetector_process("BGRx", "RGBA", w, h, d_a, 4 * w, d_b, 4 * w, **dkw)        # device -> device
            ctx.colordetect_histogram("RGBA", w, h, d_b, 4 * w, 10, d_hist)                     # device -> device
            ctx.blockhash_sums("RGBA", w, h, d_b, 4 * w, d_sums)
            ctx.download(out, 4 * w, d_b, 4 * w, 4 * w, h)
            ctx.download(hist, 4 * 32768, d_hist, 4 * 32768, 4 * 32768, 1)
            ctx.download(sums, 256, d_sums, 256, 256, 1)
            ctx.synchronize()
            assert (out == step2).all()
            assert (hist == orc.colordetect_histogram("RGBA", w, h, step2, 10)).all()
            assert (sums == orc.blockhash_sums("RGBA", w, h, step2)).all()
        for p in (d_a, d_b, d_hist, d_sums):
            ctx.device_free(p)
        with pytest.raises(b200vfx.B200VfxError):
            ctx.upload(0, 4 * w, host_in, 4 * w, 4 * w, h)


def test_hsv_zero_copy_host_path_and_cross_stream_table_use():
    """memoised hsvfilter / hsvdetector on PINNED host frames: once the answer table exists the frame is processed by one
    TMA streaming kernel that reads and writes host memory directly (in place for hsvfilter).  Also the cross-stream
    case: the table is built by a DEVICE-frame call on the caller's stream and used right away by a host-frame call on
    the library's own stream (must wait for the build)."""
    torch = pytest.importorskip("torch")
    fkw, fokw = dict(hue_shift=70.0, value_mul=0.9), dict(hue_shift=70.0, val_mul=0.9)
    dkw = dict(hue_ref=100.0, hue_var=80.0, saturation_ref=0.5, saturation_var=0.5, value_ref=0.5, value_var=0.5)
    dokw = dict(hue_ref=100.0, hue_var=80.0, sat_ref=0.5, sat_var=0.5, val_ref=0.5, val_var=0.5)
    for fmt, ofmt in (("RGBx", "RGBA"), ("xBGR", "ARGB"), ("BGRx", "ABGR")):
        for (w, h) in ((1280, 360), (644, 33), (62, 7)):           # 62: not a multiple of 4 -> staged pipeline
            with b200vfx.Context(0) as ctx:
                s = torch.cuda.Stream()
                ctx.set_stream(s.cuda_stream)
                ctx.set_option("hsv_memo", 1)
                frame = synth.frame_noise(fmt, w, h, 91 + w)
                exp_f = orc.hsvfilter(fmt, w, h, frame.copy(), threads=8, **fokw)
                exp_d = orc.hsvdetector(fmt, ofmt, w, h, frame, threads=8, **dokw)
                # table built by a device-frame call on stream s ...
                d = torch.from_numpy(frame).cuda()
                with torch.cuda.stream(s):
                    ctx.hsvfilter_process(fmt, w, h, d, frame.shape[1], **fkw)
                    dd = torch.empty((h, 4 * w), dtype=torch.uint8, device="cuda")
                    ctx.hsvdetector_process(fmt, ofmt, w, h, torch.from_numpy(frame).cuda(), frame.shape[1], dd, 4 * w, **dkw)
                # ... and used at once by host-frame calls (pinned: zero-copy kernel on the library's stream)
                for rep in range(2):
                    hp = torch.from_numpy(frame.copy()).pin_memory()
                    ctx.hsvfilter_process(fmt, w, h, hp.numpy(), frame.shape[1], **fkw)
                    assert (hp.numpy() == exp_f).all(), (fmt, w, h, rep)
                    hi = torch.from_numpy(frame.copy()).pin_memory()
                    ho = torch.full((h, 4 * w), 0x5A, dtype=torch.uint8).pin_memory()
                    ctx.hsvdetector_process(fmt, ofmt, w, h, hi.numpy(), frame.shape[1], ho.numpy(), 4 * w, **dkw)
                    assert (ho.numpy() == exp_d).all(), (fmt, ofmt, w, h, rep)
                s.synchronize()
                assert (d.cpu().numpy() == exp_f).all() and (dd.cpu().numpy() == exp_d).all()
                ctx.set_option("zero_copy", 0)                      # the staged pipeline gives the same bytes
                hp = torch.from_numpy(frame.copy()).pin_memory()
                ctx.hsvfilter_process(fmt, w, h, hp.numpy(), frame.shape[1], **fkw)
                assert (hp.numpy() == exp_f).all()
