#!/usr/bin/env python
"""Per-kernel device-resident timings (CUDA events) for every element of the hot path, with the
algorithmic-byte roofline of SURVEY 8(d) next to each.  Development tool: the numbers that are
judged come from bench.py; this prints one JSON object per line so a gpurun call can collect the
whole matrix at once.  Usage: python scripts/kernel_bench.py [--only colorlut,hsv,...] [--iters 50]"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gst-plugin-rs_b200"))

import numpy as np
import torch

import b200vfx
from b200vfx import synth


def peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


PEAK = peak()
RING = 6  # distinct in/out buffers so consecutive launches do not hit L2-resident frames


def timeit(fn, iters, warm=5):
    for _ in range(warm):
        fn(0)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(iters):
        fn(i)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e-3 / iters


def graph_time(fn, launches=16, replays=30):
    """device time per launch without the host enqueue: `launches` calls captured into one CUDA graph, replayed"""
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        for i in range(3):
            fn(i)
    st.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=st):
        for i in range(launches):
            fn(i)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(replays):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e-3 / (replays * launches)


def report(name, seconds, algo_bytes, **extra):
    gbs = algo_bytes / seconds / 1e9
    print(json.dumps({"kernel": name, "us": round(seconds * 1e6, 3), "algo_GBps": round(gbs, 1), "frac_of_measured_peak": round(gbs / PEAK, 4),
                      **extra}), flush=True)


def ring_of(frame_fn, n=RING):
    frames = [torch.from_numpy(frame_fn(i)).cuda() for i in range(n)]
    outs = [torch.empty_like(f) for f in frames]
    return frames, outs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    ap.add_argument("--iters", type=int, default=60)
    args = ap.parse_args()
    only = set(filter(None, args.only.split(",")))
    want = lambda k: not only or k in only
    ctx = b200vfx.Context(0)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    W, H = 3840, 2160
    contents = {
        "ramps": lambda i: np.ascontiguousarray(np.roll(synth.frame_ramps("RGBA", W, H), 4 * 131 * i, axis=1)),
        "noise": lambda i: synth.frame_noise("RGBA", W, H, 100 + i),
        "natural": lambda i: synth.frame_natural("RGBA", W, H, 200 + i, amp=3),
    }
    if want("colorlut"):
        for n in (33, 65):
            k, s, v, sc, of = b200vfx.cube_parse(synth.cube_text_3d(n, "mix"))
            ctx.colorlut_set_lut(k, s, v, sc, of)
            for cname, fn in contents.items():
                frames, outs = ring_of(fn)
                for mode in (0, 1):
                    ctx.colorlut_set_mode(mode)
                    t = timeit(lambda i: ctx.colorlut_process("RGBA", W, H, frames[i % RING], 4 * W, outs[i % RING], 4 * W), args.iters)
                    report("colorlut3d_%s_rgba8" % ("memo" if mode == 0 else "direct"), t, 2 * W * H * 4, lut=n, content=cname, frame="3840x2160")
                del frames, outs
            ctx.colorlut_set_mode(0)
        # 1D LUT
        k, s, v, sc, of = b200vfx.cube_parse(synth.cube_text_1d(1024, 2.2))
        ctx.colorlut_set_lut(k, s, v, sc, of)
        for cname in ("ramps", "noise"):
            frames, outs = ring_of(contents[cname])
            for mode in (0, 1):
                ctx.colorlut_set_mode(mode)
                t = timeit(lambda i: ctx.colorlut_process("RGBA", W, H, frames[i % RING], 4 * W, outs[i % RING], 4 * W), args.iters)
                report("colorlut1d_%s_rgba8" % ("memo" if mode == 0 else "direct"), t, 2 * W * H * 4, lut=1024, content=cname, frame="3840x2160")
        ctx.colorlut_set_mode(0)
    if want("colorlut64"):
        k, s, v, sc, of = b200vfx.cube_parse(synth.cube_text_3d(33, "mix"))
        ctx.colorlut_set_lut(k, s, v, sc, of)
        for fmt in ("RGBA64_LE", "RGBA64_BE"):
            for cname, fn in (("ramps", lambda i: np.ascontiguousarray(np.roll(synth.frame_ramps(fmt, W, H), 8 * 131 * i, axis=1))),
                              ("noise", lambda i: synth.frame_noise(fmt, W, H, 300 + i))):
                frames, outs = ring_of(fn, 4)
                t = timeit(lambda i: ctx.colorlut_process(fmt, W, H, frames[i % 4], 8 * W, outs[i % 4], 8 * W), max(args.iters // 3, 5))
                report("colorlut3d_direct_%s" % fmt.lower(), t, 2 * W * H * 8, lut=33, content=cname, frame="3840x2160")
                del frames, outs
    if want("hsv"):
        kw = dict(hue_ref=120.0, hue_var=30.0, saturation_ref=0.8, saturation_var=0.2, value_ref=0.8, value_var=0.2)
        for memo in (0, 1):
            ctx.set_option("hsv_memo", memo)
            tag_m = "memo" if memo else "direct"
            for (w, h, tag) in ((640, 480, "640x480"), (3840, 2160, "3840x2160")):
                for cname in ("ramps", "noise"):
                    fr = [torch.from_numpy(synth.frame_ramps("RGBA", w, h) if cname == "ramps" else synth.frame_noise("RGBA", w, h, 0x5EED0001 + i)).cuda() for i in range(RING)]
                    t = timeit(lambda i: ctx.hsvfilter_process("RGBA", w, h, fr[i % RING], 4 * w, hue_shift=90.0), args.iters)
                    report("hsvfilter_rgba_" + tag_m, t, 2 * w * h * 4, content=cname, frame=tag)
            w, h = 1920, 1080
            for cname in ("ramps", "noise"):
                fr = [torch.from_numpy(synth.frame_ramps("BGRx", w, h) if cname == "ramps" else synth.frame_noise("BGRx", w, h, 0x5EED0003 + i)).cuda() for i in range(RING)]
                out = [torch.empty_like(f) for f in fr]
                t = timeit(lambda i: ctx.hsvdetector_process("BGRx", "RGBA", w, h, fr[i % RING], 4 * w, out[i % RING], 4 * w, **kw), args.iters)
                report("hsvdetector_bgrx_rgba_" + tag_m, t, 2 * w * h * 4, content=cname, frame="1920x1080")
            for (w, h) in ((1920, 1080), (3840, 2160)):
                fr = [torch.from_numpy(synth.frame_noise("RGB", w, h, 9 + i)).cuda() for i in range(4)]
                out = [torch.empty((h, 4 * w), dtype=torch.uint8, device="cuda") for _ in range(4)]
                t = timeit(lambda i: ctx.hsvdetector_process("RGB", "ARGB", w, h, fr[i % 4], 3 * w, out[i % 4], 4 * w, **kw), max(args.iters // 3, 5))
                report("hsvdetector_rgb_argb_" + tag_m, t, w * h * 7, content="noise", frame="%dx%d" % (w, h))
        ctx.set_option("hsv_memo", -1)
    if want("hsvanim"):
        # animated properties (a GstController changing hue-shift on every frame): no answer table survives, every frame
        # takes the direct kernel -- default policy (hsv_memo = -1), per-frame distinct settings
        ctx.set_option("hsv_memo", -1)
        for cname in ("ramps", "noise", "natural"):
            fr = [torch.from_numpy(contents[cname](i)).cuda() for i in range(RING)]
            t = timeit(lambda i: ctx.hsvfilter_process("RGBA", W, H, fr[i % RING], 4 * W, hue_shift=0.37 * (i % 900) - 120.0, saturation_mul=1.1), args.iters)
            report("hsvfilter_rgba_animated_hue_shift", t, 2 * W * H * 4, content=cname, frame="3840x2160")
            out = [torch.empty_like(f) for f in fr]
            t = timeit(lambda i: ctx.hsvdetector_process("BGRx", "RGBA", W, H, fr[i % RING], 4 * W, out[i % RING], 4 * W, hue_ref=0.5 * (i % 700), hue_var=30.0,
                                                         saturation_ref=0.8, saturation_var=0.2, value_ref=0.8, value_var=0.2), args.iters)
            report("hsvdetector_bgrx_rgba_animated_hue_ref", t, 2 * W * H * 4, content=cname, frame="3840x2160")
            del fr, out
    if want("videofx"):
        frames, _ = ring_of(contents["noise"])
        sums = torch.zeros(64, dtype=torch.int32, device="cuda")
        t = timeit(lambda i: ctx.blockhash_sums("RGBA", W, H, frames[i % RING], 4 * W, sums), args.iters)
        report("blockhash_sums_rgba", t, W * H * 4, content="noise", frame="3840x2160", note="one stream; config 4 = two streams; per-call API loop (host enqueue included)")
        try:   # the same launches replayed from a CUDA graph: device time only
            cur = torch.cuda.current_stream()
            def in_graph(fn):
                def run(i):
                    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
                    fn(i)
                return run
            t = graph_time(in_graph(lambda i: ctx.blockhash_sums("RGBA", W, H, frames[i % RING], 4 * W, sums)))
            report("blockhash_sums_rgba", t, W * H * 4, content="noise", frame="3840x2160", note="CUDA-graph replay: device time per launch")
            sums2g = torch.zeros(128, dtype=torch.int32, device="cuda")
            t = graph_time(in_graph(lambda i: ctx.blockhash_sums_batch("RGBA", W, H, [frames[(2 * i) % RING], frames[(2 * i + 1) % RING]], [4 * W, 4 * W], sums2g)))
            report("blockhash_sums_rgba_batch2", t, 2 * W * H * 4, content="noise", frame="2 x 3840x2160", note="CUDA-graph replay: device time per launch")
            histg = torch.zeros(32768, dtype=torch.int32, device="cuda")
            for q in (10, 1):
                t = graph_time(in_graph(lambda i: ctx.colordetect_histogram("RGBA", W, H, frames[i % RING], 4 * W, q, histg)))
                report("colordetect_hist_rgba", t, W * H * 4, content="noise", frame="3840x2160", quality=q, note="CUDA-graph replay: device time per launch")
            ctx.set_stream(cur.cuda_stream)
        except Exception as exc:
            print(json.dumps({"graph_timing_error": str(exc)[:200]}), flush=True)
            ctx.set_stream(torch.cuda.current_stream().cuda_stream)
        sums2 = torch.zeros(128, dtype=torch.int32, device="cuda")
        t = timeit(lambda i: ctx.blockhash_sums_batch("RGBA", W, H, [frames[(2 * i) % RING], frames[(2 * i + 1) % RING]], [4 * W, 4 * W], sums2), args.iters)
        report("blockhash_sums_rgba_batch2", t, 2 * W * H * 4, content="noise", frame="2 x 3840x2160", note="BASELINE config 4: both streams in one launch")
        m = torch.empty((1080, 1920), dtype=torch.uint8, device="cuda")
        t = timeit(lambda i: ctx.roundmask_generate(1920, 1080, 1920, 64, m), 20)
        report("roundmask_a8", t, 1920 * 1080, frame="1920x1080", radius=64, note="once per caps/radius change")
    if want("planar"):
        # colorlut on I420 frames with both videoconverts fused: 1.5 B/px in, 1.5 B/px out
        k, s_, v, sc, of = b200vfx.cube_parse(synth.cube_text_3d(33, "mix"))
        ctx.colorlut_set_lut(k, s_, v, sc, of)
        rng = np.random.default_rng(5)
        strides = [W, W // 2, W // 2]
        shapes = [(H, W), (H // 2, W // 2), (H // 2, W // 2)]
        yy, xx = np.mgrid[0:H, 0:W]
        smooth = [((xx * 200 // W) + (yy * 40 // H) + 16).astype(np.uint8), np.full(shapes[1], 110, np.uint8), np.full(shapes[2], 150, np.uint8)]
        gens = {"noise": lambda i: [rng.integers(0, 256, sh, dtype=np.uint8) for sh in shapes],
                "natural": lambda i: [np.clip(b.astype(np.int16) + rng.integers(-3, 4, b.shape), 0, 255).astype(np.uint8) for b in smooth]}
        for cname, gen in gens.items():
            src = [[torch.from_numpy(p).cuda() for p in gen(i)] for i in range(RING)]
            dst = [[torch.empty_like(p) for p in f] for f in src]
            t = timeit(lambda i: ctx.colorlut_process_planar("I420", W, H, src[i % RING], strides, dst[i % RING], strides), args.iters)
            report("colorlut_i420_fused_converts", t, W * H * 3, content=cname, frame="3840x2160", note="I420 -> RGB -> LUT 33^3 -> I420 in one kernel")
            rgba = [torch.empty((H, 4 * W), dtype=torch.uint8, device="cuda") for _ in range(2)]
            def chain(i):
                ctx.convert_from_planar("I420", "RGBA", W, H, src[i % RING], strides, rgba[0], 4 * W)
                ctx.colorlut_process("RGBA", W, H, rgba[0], 4 * W, rgba[1], 4 * W)
                ctx.convert_to_planar("RGBA", "I420", W, H, rgba[1], 4 * W, dst[i % RING], strides)
            t = timeit(chain, args.iters)
            report("colorlut_i420_three_kernels", t, W * H * 3, content=cname, frame="3840x2160", note="convert_from_planar + colorlut + convert_to_planar")
            del src, dst, rgba
    if want("hashes"):
        # videocompare's other hash algorithms: grayscale + Lanczos3 resize to 8x8 / 9x8 / 8x9 / 5x5 (two kernels) + the
        # 64-81 byte read-back the bit rule needs on the host -- a synchronous call, timed by the host clock
        import time
        frames, _ = ring_of(contents["natural"])
        for algo in ("mean", "gradient", "vertgradient", "doublegradient", "blockhash"):
            for i in range(4):
                ctx.hash_image(algo, "RGBA", W, H, frames[i % RING], 4 * W)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for i in range(args.iters):
                ctx.hash_image(algo, "RGBA", W, H, frames[i % RING], 4 * W)
            t = (time.perf_counter() - t0) / args.iters
            report("videocompare_hash_image_" + algo, t, W * H * 4, content="natural", frame="3840x2160", note="synchronous call on a device frame: kernels + read-back + host bit rule")
    if want("colordetect"):
        hist = torch.zeros(32768, dtype=torch.int32, device="cuda")
        for cname in ("ramps", "noise", "natural"):
            frames, _ = ring_of(contents[cname])
            for q, cl in ((10, 2), (1, 2), (1, 1), (1, 4), (1, 8), (10, 1)):
                ctx.set_option("cd_cluster", cl)
                t = timeit(lambda i: ctx.colordetect_histogram("RGBA", W, H, frames[i % RING], 4 * W, q, hist), args.iters)
                # quality q reads one 4-byte pixel every 4q bytes: every 32-byte sector is still touched up to q = 8
                report("colordetect_hist_rgba", t, W * H * 4, content=cname, frame="3840x2160", quality=q, cluster=cl,
                       samples=(W * H + q - 1) // q, note="algorithmic bytes = the whole plane (sector granularity)")
            del frames
        ctx.set_option("cd_cluster", 2)
    if want("hsv24"):
        for memo in (0, 1):
            ctx.set_option("hsv_memo", memo)
            for cname in ("ramps", "noise"):
                fr = [torch.from_numpy((synth.frame_ramps("RGB", W, H) if cname == "ramps" else synth.frame_noise("RGB", W, H, 50 + i))[:, :3 * W].copy()).cuda() for i in range(RING)]
                t = timeit(lambda i: ctx.hsvfilter_process("RGB", W, H, fr[i % RING], 3 * W, hue_shift=90.0), max(args.iters // (1 if memo else 3), 5))
                report("hsvfilter_rgb_" + ("memo" if memo else "direct"), t, 2 * W * H * 3, content=cname, frame="3840x2160")
        ctx.set_option("hsv_memo", -1)
    if want("e2e"):
        k, s, v, sc, of = b200vfx.cube_parse(synth.cube_text_3d(33, "mix"))
        ctx.colorlut_set_lut(k, s, v, sc, of)
        src = torch.from_numpy(contents["noise"](0)).pin_memory()
        dst = torch.empty_like(src).pin_memory()
        pag_src = contents["noise"](1)
        pag_dst = np.empty_like(pag_src)
        ctx.set_option("zero_copy", 0)   # first the staged (copy-engine) pipeline, chunk size swept
        for rows in (0, 136, 270, 540, 2160):
            ctx.set_chunk_rows(rows)
            ctx.colorlut_process("RGBA", W, H, src.numpy(), 4 * W, dst.numpy(), 4 * W)
            t0 = time.perf_counter()
            n = 12
            for _ in range(n):
                ctx.colorlut_process("RGBA", W, H, src.numpy(), 4 * W, dst.numpy(), 4 * W)
            t = (time.perf_counter() - t0) / n
            print(json.dumps({"kernel": "e2e_colorlut_pinned", "chunk_rows": rows, "ms": round(t * 1e3, 4), "fps": round(1 / t, 1),
                              "pcie_GBps_each_way": round(W * H * 4 / t / 1e9, 2)}), flush=True)
        ctx.set_chunk_rows(0)
        ctx.set_option("zero_copy", 0)
        ctx.set_chunk_rows(540)
        ctx.colorlut_process("RGBA", W, H, src.numpy(), 4 * W, dst.numpy(), 4 * W)
        ctx.set_chunk_rows(0)
        ref_out = dst.clone()
        for cfg, ctas, grid in ((2, 1, 0), (2, 1, 64), (2, 1, 32), (2, 1, 16), (2, 2, 0), (7, 1, 0), (7, 2, 0), (7, 1, 64), (4, 1, 0), (4, 2, 0),
                                (1, 1, 0), (1, 1, 64), (0, 1, 0), (0, 1, 32)):
                ctx.set_option("zero_copy", 1); ctx.set_option("zc_cfg", cfg); ctx.set_option("zc_ctas", ctas); ctx.set_option("zc_grid", grid)
                dst.zero_()
                ctx.colorlut_process("RGBA", W, H, src.numpy(), 4 * W, dst.numpy(), 4 * W)
                same = bool((dst == ref_out).all())
                t0 = time.perf_counter()
                for _ in range(12):
                    ctx.colorlut_process("RGBA", W, H, src.numpy(), 4 * W, dst.numpy(), 4 * W)
                t = (time.perf_counter() - t0) / 12
                print(json.dumps({"kernel": "e2e_colorlut_pinned_zero_copy", "stream_cfg": cfg, "ctas_per_sm": ctas, "grid_cap": grid, "ms": round(t * 1e3, 4),
                                  "fps": round(1 / t, 1), "pcie_GBps_each_way": round(W * H * 4 / t / 1e9, 2), "identical": same}), flush=True)
        ctx.set_option("zero_copy", 0); ctx.set_option("stream_cfg", 0); ctx.set_option("stream_ctas", 0)
        ctx.colorlut_process("RGBA", W, H, pag_src, 4 * W, pag_dst, 4 * W)
        t0 = time.perf_counter()
        for _ in range(6):
            ctx.colorlut_process("RGBA", W, H, pag_src, 4 * W, pag_dst, 4 * W)
        t = (time.perf_counter() - t0) / 6
        print(json.dumps({"kernel": "e2e_colorlut_pageable", "ms": round(t * 1e3, 4), "fps": round(1 / t, 1)}), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
