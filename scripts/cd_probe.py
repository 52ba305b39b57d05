import sys, os, json, time
sys.path.insert(0, "gst-plugin-rs_b200")
import numpy as np, torch, b200vfx
from b200vfx import synth
ctx = b200vfx.Context(0); ctx.set_stream(torch.cuda.current_stream().cuda_stream)
hist = torch.zeros(32768, dtype=torch.int32, device="cuda")
def timeit(fn, n=200):
    for _ in range(10): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); a.record()
    for _ in range(n): fn()
    b.record(); t1 = time.perf_counter(); torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / n, (t1 - t0) * 1e6 / n
for (w, h) in ((64, 64), (640, 480), (1920, 1080), (3840, 2160)):
    fr = torch.from_numpy(synth.frame_noise("RGBA", w, h, 1)).cuda()
    for q in (10, 1):
        dev, host = timeit(lambda: ctx.colordetect_histogram("RGBA", w, h, fr, 4 * w, q, hist))
        print(json.dumps({"frame": "%dx%d" % (w, h), "q": q, "dev_us": round(dev, 2), "host_enqueue_us": round(host, 2)}))
