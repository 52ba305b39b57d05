#!/usr/bin/env python
"""One line of JSON per BASELINE.json config (1-5): device-resident time, end-to-end (pinned host buffers) rate and the
CPU oracle port beside it.  Config 5 (8K frame row-tiled over all ranks + NCCL all-gather) needs torchrun:
    python scripts/config_bench.py                       # configs 1-4 on one GPU
    torchrun --nproc-per-node 8 scripts/config_bench.py --config5
The judged headline comes from bench.py; this is the per-config companion table (profiles/r01_configs.jsonl)."""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gst-plugin-rs_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)

import numpy as np
import torch

import b200vfx
from b200vfx import sharding, synth
from bench import host_threads, hbm_peak_gbs

PEAK = hbm_peak_gbs()[0]


def dev_time(fn, iters=50, warm=5):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(iters):
        fn(i)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e-3 / iters


def wall_time(fn, iters=12, warm=7):
    for i in range(warm):
        fn(i)
    t0 = time.perf_counter()
    for i in range(iters):
        fn(i)
    return (time.perf_counter() - t0) / iters


def async_wall_time(ctx, fn, iters, warm, inflight=3):
    """the same calls in asynchronous host-frame mode: <= `inflight` frames in flight (a fence per frame), final synchronise
    inside the timed region; the callers' rings hold more buffers than that"""
    for i in range(warm):
        fn(i)
    ctx.set_host_async(True)
    fences = []
    for i in range(4):
        fn(i); fences.append(ctx.fence())
    ctx.synchronize()
    t0 = time.perf_counter()
    for i in range(iters):
        if len(fences) >= inflight:
            f = fences.pop(0); f.wait(); f.close()
        fn(i)
        fences.append(ctx.fence())
    ctx.synchronize()
    dt = (time.perf_counter() - t0) / iters
    for f in fences:
        f.close()
    ctx.set_host_async(False)
    return dt


def cpu_time(fn, budget=4.0):
    fn()
    n, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < budget:
        fn()
        n += 1
    return (time.perf_counter() - t0) / n


def pin(a):
    return torch.from_numpy(a).pin_memory()


def configs_1_to_4():
    import oracle_binding as orc
    T = host_threads()
    ctx = b200vfx.Context(0)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    R = 6
    # ---- config 1: hsvfilter hue-shift on 640x480 RGBA ----------------------------------------------------------
    w, h = 640, 480
    frames = [synth.frame_ramps("RGBA", w, h) if i % 2 == 0 else synth.frame_noise("RGBA", w, h, 0x5EED0001 + i) for i in range(R)]
    d = [torch.from_numpy(f).cuda() for f in frames]
    hp = [pin(f) for f in frames]
    # warm-up long enough for the rent-or-buy policy to have built the memo table (2^24 px with unchanged settings):
    # the numbers below are the steady state of a stream, the one-off table build is reported separately
    warm = (1 << 24) // (w * h) + 4
    t_dev = dev_time(lambda i: ctx.hsvfilter_process("RGBA", w, h, d[i % R], 4 * w, hue_shift=90.0), 200, warm)
    t_e2e = wall_time(lambda i: ctx.hsvfilter_process("RGBA", w, h, hp[i % R].numpy(), 4 * w, hue_shift=90.0), 200, 60)
    t_e2e_async = async_wall_time(ctx, lambda i: ctx.hsvfilter_process("RGBA", w, h, hp[i % R].numpy(), 4 * w, hue_shift=90.0), 400, 10)
    ctx.set_option("hsv_memo", 0)
    t_dev_direct = dev_time(lambda i: ctx.hsvfilter_process("RGBA", w, h, d[i % R], 4 * w, hue_shift=90.0), 200, 5)
    ctx.set_option("hsv_memo", -1)
    t_cpu1 = cpu_time(lambda: orc.hsvfilter("RGBA", w, h, frames[0], hue_shift=90.0, threads=1), 1.5)
    t_cpuN = cpu_time(lambda: orc.hsvfilter("RGBA", w, h, frames[0], hue_shift=90.0, threads=T), 1.5)
    print(json.dumps({"config": 1, "what": "hsvfilter hue-shift=90, 640x480 RGBA, in place", "device_us": round(t_dev * 1e6, 2),
                      "device_fps": round(1 / t_dev), "algo_GBps": round(2 * w * h * 4 / t_dev / 1e9, 1), "device_us_direct_kernel": round(t_dev_direct * 1e6, 2),
                      "e2e_fps": round(1 / t_e2e), "e2e_fps_async_mode": round(1 / t_e2e_async), "cpu_fps_1thread": round(1 / t_cpu1, 1), "cpu_fps_%dthreads" % T: round(1 / t_cpuN, 1)}), flush=True)
    # ---- config 3: hsvdetector (BGRx -> RGBA) + roundedcorners mask, 1920x1080 -------------------------------------
    w, h = 1920, 1080
    kw = dict(hue_ref=120.0, hue_var=30.0, saturation_ref=0.8, saturation_var=0.2, value_ref=0.8, value_var=0.2)
    okw = dict(hue_ref=120.0, hue_var=30.0, sat_ref=0.8, sat_var=0.2, val_ref=0.8, val_var=0.2)
    frames = [synth.frame_ramps("BGRx", w, h) if i % 2 == 0 else synth.frame_noise("BGRx", w, h, 0x5EED0003 + i) for i in range(R)]
    d = [torch.from_numpy(f).cuda() for f in frames]
    do = [torch.empty_like(x) for x in d]
    hp, ho = [pin(f) for f in frames], [pin(np.zeros_like(f)) for f in frames]
    warm = (1 << 24) // (w * h) + 4
    t_dev = dev_time(lambda i: ctx.hsvdetector_process("BGRx", "RGBA", w, h, d[i % R], 4 * w, do[i % R], 4 * w, **kw), 100, warm)
    t_e2e = wall_time(lambda i: ctx.hsvdetector_process("BGRx", "RGBA", w, h, hp[i % R].numpy(), 4 * w, ho[i % R].numpy(), 4 * w, **kw), 60, 20)
    t_e2e_async = async_wall_time(ctx, lambda i: ctx.hsvdetector_process("BGRx", "RGBA", w, h, hp[i % R].numpy(), 4 * w, ho[i % R].numpy(), 4 * w, **kw), 120, 6)
    ctx.set_option("hsv_memo", 0)
    t_dev_direct = dev_time(lambda i: ctx.hsvdetector_process("BGRx", "RGBA", w, h, d[i % R], 4 * w, do[i % R], 4 * w, **kw), 100, 5)
    ctx.set_option("hsv_memo", -1)
    t_cpu1 = cpu_time(lambda: orc.hsvdetector("BGRx", "RGBA", w, h, frames[1], threads=1, **okw), 2.0)
    t_cpuN = cpu_time(lambda: orc.hsvdetector("BGRx", "RGBA", w, h, frames[1], threads=T, **okw), 2.0)
    mask = torch.empty((h, w), dtype=torch.uint8, device="cuda")
    t_mask = dev_time(lambda i: ctx.roundmask_generate(w, h, w, 64, mask), 20)
    t_mask_cpu = cpu_time(lambda: orc.roundmask(w, h, w, 64), 1.0)
    print(json.dumps({"config": 3, "what": "hsvdetector BGRx->RGBA 1920x1080 (+ roundedcorners r=64 mask once per caps/radius; the two "
                      "elements cannot be linked directly: SURVEY D1)", "device_us": round(t_dev * 1e6, 2), "device_fps": round(1 / t_dev),
                      "algo_GBps": round(2 * w * h * 4 / t_dev / 1e9, 1), "device_us_direct_kernel": round(t_dev_direct * 1e6, 2), "e2e_fps": round(1 / t_e2e), "e2e_fps_async_mode": round(1 / t_e2e_async),
                      "cpu_fps_1thread": round(1 / t_cpu1, 1), "cpu_fps_%dthreads" % T: round(1 / t_cpuN, 1), "mask_device_us": round(t_mask * 1e6, 1), "mask_cpu_us": round(t_mask_cpu * 1e6, 1)}), flush=True)
    # ---- chain: hsvfilter -> hsvdetector on 1920x1080 BGRx, host frame in, host frame out -------------
    # (a) as two unchanged elements with system memory between them: every element uploads and downloads its frame;
    # (b) device-resident: one upload (b200vfx_upload), both kernels on the frame in HBM, one download
    d_a, d_b = ctx.device_alloc(4 * w * h), ctx.device_alloc(4 * w * h)
    mid = [pin(np.zeros_like(f)) for f in frames]

    def chain_host(i):
        mid[i % R].copy_(hp[i % R])                                                     # hsvfilter works in place on its input buffer
        ctx.hsvfilter_process("BGRx", w, h, mid[i % R].numpy(), 4 * w, hue_shift=90.0)
        ctx.hsvdetector_process("BGRx", "RGBA", w, h, mid[i % R].numpy(), 4 * w, ho[i % R].numpy(), 4 * w, **kw)

    def chain_device(i):
        ctx.upload(d_a, 4 * w, hp[i % R].numpy(), 4 * w, 4 * w, h)
        ctx.hsvfilter_process("BGRx", w, h, d_a, 4 * w, hue_shift=90.0)
        ctx.hsvdetector_process("BGRx", "RGBA", w, h, d_a, 4 * w, d_b, 4 * w, **kw)
        ctx.download(ho[i % R].numpy(), 4 * w, d_b, 4 * w, 4 * w, h)
        ctx.synchronize()

    warm = 2 * ((1 << 24) // (w * h) + 4)
    t_host = wall_time(chain_host, 60, warm)
    ref_out = ho[0].clone()
    t_devc = wall_time(chain_device, 60, warm)
    same = bool((ho[0] == ref_out).all())
    print(json.dumps({"config": "3-chain", "what": "hsvfilter (in place) -> hsvdetector BGRx->RGBA, 1920x1080, pinned host frame in and out",
                      "two_host_elements_fps": round(1 / t_host), "device_resident_chain_fps": round(1 / t_devc), "identical": same,
                      "note": "device-resident: one upload + two kernels + one download per frame (b200vfx_upload / _download)"}), flush=True)
    ctx.device_free(d_a); ctx.device_free(d_b)
    # ---- config 4: videocompare blockhash on two 3840x2160 RGBA streams ---------------------------------------------
    w, h = 3840, 2160
    a = synth.frame_ramps("RGBA", w, h)
    b = a.copy()
    idx = synth.pcg32(w * h // 100, 0x5EED0004) % np.uint32(w * h)
    b.reshape(-1, 4)[idx, :3] ^= 0x80
    da, db = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    sa, sb = torch.zeros(64, dtype=torch.int32, device="cuda"), torch.zeros(64, dtype=torch.int32, device="cuda")
    pa, pb = pin(a), pin(b)
    ha, hb = np.zeros(64, np.uint32), np.zeros(64, np.uint32)

    sab = torch.zeros(128, dtype=torch.int32, device="cuda")
    hab = np.zeros(128, np.uint32)
    # more than one pair of device frames so that consecutive compares do not find their inputs in L2
    ring = [(da, db)] + [(da.clone(), db.clone()) for _ in range(2)]

    def compare_dev(i):      # what the element does: both pads' frames in ONE launch (b200vfx_blockhash_sums_batch)
        x, y = ring[i % 3]
        ctx.blockhash_sums_batch("RGBA", w, h, [x, y], [4 * w, 4 * w], sab)

    def compare_e2e(i):
        ctx.blockhash_sums_batch("RGBA", w, h, [pa.numpy(), pb.numpy()], [4 * w, 4 * w], hab)
        ha[:], hb[:] = hab[:64], hab[64:]
        return b200vfx.hash_distance(b200vfx.blockhash_bits(ha, w, h), b200vfx.blockhash_bits(hb, w, h))

    t_dev = dev_time(compare_dev, 60)
    t_e2e = wall_time(compare_e2e, 20, 5)
    dist = compare_e2e(0)

    t_cpu = cpu_time(lambda: (orc.blockhash_sums("RGBA", w, h, a), orc.blockhash_sums("RGBA", w, h, b)), 3.0)
    print(json.dumps({"config": 4, "what": "videocompare blockhash, two 3840x2160 RGBA streams (stream 1 = stream 0 with 1% perturbed pixels)",
                      "device_us_per_compare": round(t_dev * 1e6, 2), "device_compares_per_s": round(1 / t_dev),
                      "algo_GBps": round(2 * w * h * 4 / t_dev / 1e9, 1), "frac_of_measured_peak": round(2 * w * h * 4 / t_dev / 1e9 / PEAK, 3),
                      "e2e_compares_per_s": round(1 / t_e2e, 1), "hamming_distance": dist, "cpu_compares_per_s_1thread": round(1 / t_cpu, 2)}), flush=True)
    ctx.close()


def config5():
    import torch.distributed as dist
    import oracle_binding as orc
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    W, H, R = 7680, 4320, 4
    ctx = b200vfx.Context(local)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    k, s, v, sc, of = b200vfx.cube_parse(synth.cube_text_3d(65, "mix"))
    ctx.colorlut_set_lut(k, s, v, sc, of)
    r0, r1 = sharding.row_range(H, world, rank)
    frames = [synth.frame_ramps("RGBA", W, H)[r0:r1].copy() if i % 2 == 0 else synth.frame_noise("RGBA", W, r1 - r0, 0x5EED0005 + 16 * rank + i) for i in range(R)]
    d_in = [torch.from_numpy(f).cuda() for f in frames]
    d_out = [torch.empty_like(x) for x in d_in]
    rows = r1 - r0

    def kern(i):
        ctx.colorlut_process("RGBA", W, rows, d_in[i % R], 4 * W, d_out[i % R], 4 * W)

    def kern_gather(i):
        kern(i)
        sharding.all_gather_rows(dist, d_out[i % R], H, world)

    dist.barrier()
    t_k = dev_time(kern, 100)
    dist.barrier()
    t_kg = dev_time(kern_gather, 50)
    t_g = dev_time(lambda i: sharding.all_gather_rows(dist, d_out[i % R], H, world), 50)
    tt = torch.tensor([t_k, t_kg, t_g], dtype=torch.float64, device="cuda")
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_k, t_kg, t_g = [float(x) for x in tt.tolist()]
    if rank == 0:
        print(json.dumps({"config": 5, "what": "colorlut 65^3 on 7680x4320 RGBA row-tiled over %d B200s (%d rows per GPU), frames A/B alternating" % (world, rows),
                          "tile_kernel_us": round(t_k * 1e6, 2), "frames_per_s_no_gather": round(1 / t_k),
                          "per_gpu_algo_GBps": round(2 * W * rows * 4 / t_k / 1e9, 1), "per_gpu_frac_of_measured_peak": round(2 * W * rows * 4 / t_k / 1e9 / PEAK, 3),
                          "allgather_us": round(t_g * 1e6, 1), "allgather_payload_bytes_per_rank": W * rows * 4,
                          "allgather_recv_GBps_per_gpu": round(W * rows * 4 * (world - 1) / t_g / 1e9, 1),
                          "kernel_plus_allgather_us": round(t_kg * 1e6, 1), "frames_per_s_with_gather": round(1 / t_kg)}), flush=True)
    ctx.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--config5", action="store_true")
    a = ap.parse_args()
    if a.config5:
        config5()
    else:
        configs_1_to_4()
