#!/usr/bin/env python
"""The A/B stream slows down over consecutive runs inside one process (22.2 -> 24.1 us per frame) and starts fast again with
a fresh context: power / thermal state of the GPU, or state of the library?  Same stream repeated with NVML clock / power
samples, idle gaps, context re-creation, and explicit resets of the launch-overlap record."""
import json, os, sys, time, threading
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gst-plugin-rs_b200"))
import numpy as np, torch, b200vfx, pynvml
from b200vfx import synth
pynvml.nvmlInit(); hnd = pynvml.nvmlDeviceGetHandleByIndex(0)
W, H, R = 3840, 2160, 12
ramps = lambda i: np.ascontiguousarray(np.roll(synth.frame_ramps("RGBA", W, H), 4 * 97 * i, axis=1))
noise = lambda i: synth.frame_noise("RGBA", W, H, 100 + i)
host = [ramps(i // 2) if i % 2 == 0 else noise(i // 2) for i in range(R)]
k, s, v, sc, of = b200vfx.cube_parse(synth.cube_text_3d(33, "mix"))
fr = [torch.from_numpy(f).cuda() for f in host]; out = [torch.empty_like(f) for f in fr]

def run(ctx, n=64 * 20):
    samples = []
    stop = threading.Event()
    def sampler():
        while not stop.is_set():
            samples.append((pynvml.nvmlDeviceGetClockInfo(hnd, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetClockInfo(hnd, pynvml.NVML_CLOCK_MEM),
                            pynvml.nvmlDeviceGetPowerUsage(hnd) // 1000, pynvml.nvmlDeviceGetTemperature(hnd, 0)))
            time.sleep(0.002)
    th = threading.Thread(target=sampler); 
    torch.cuda.synchronize(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    th.start(); a.record()
    for i in range(n): ctx.colorlut_process("RGBA", W, H, fr[i % R], 4 * W, out[i % R], 4 * W)
    b.record(); torch.cuda.synchronize(); stop.set(); th.join()
    sm = sorted(x[0] for x in samples); 
    return {"us": round(a.elapsed_time(b) * 1e3 / n, 2), "sm_mhz_med": sm[len(sm) // 2], "sm_mhz_min": sm[0], "mem_mhz": samples[-1][1],
            "power_w_max": max(x[2] for x in samples), "temp_c": samples[-1][3], "n": len(samples)}

ctx = b200vfx.Context(0); ctx.set_stream(torch.cuda.current_stream().cuda_stream); ctx.colorlut_set_lut(k, s, v, sc, of)
print(json.dumps({"phase": "8 runs back to back", "runs": [run(ctx) for _ in range(8)]}), flush=True)
time.sleep(2.0)
print(json.dumps({"phase": "after 2 s idle", "runs": [run(ctx) for _ in range(3)]}), flush=True)
ctx.synchronize()            # our own synchronisation forgets the launch-overlap record
print(json.dumps({"phase": "after ctx.synchronize()", "runs": [run(ctx) for _ in range(3)]}), flush=True)
ctx.close()
ctx = b200vfx.Context(0); ctx.set_stream(torch.cuda.current_stream().cuda_stream); ctx.colorlut_set_lut(k, s, v, sc, of)
print(json.dumps({"phase": "fresh context, no idle", "runs": [run(ctx) for _ in range(3)]}), flush=True)
print(json.dumps({"phase": "long run 64*100", "runs": [run(ctx, 64 * 100)]}), flush=True)
ctx.close()
