#!/usr/bin/env python
"""PCIe ceiling for the e2e number: cudaMemcpyAsync of one 4K RGBA frame (33 MB) host->device, device->host, and both at
once on two streams (pinned memory), then the e2e colorlut call (zero-copy TMA kernel and staged copy-engine pipeline).
One JSON object per line.  (Write-combined pinned input memory was tried once: same PCIe rates, 6x slower CPU fill --
not offered by the library.)"""
import ctypes as C, json, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gst-plugin-rs_b200"))
import numpy as np, torch, b200vfx
from b200vfx import synth
W, H = 3840, 2160
NB = W * H * 4
L = b200vfx.lib()
rt = C.CDLL("libcudart.so.12") if False else None

def host_buf(wc):
    p = L.b200vfx_host_alloc(NB)
    assert p
    return p, np.ctypeslib.as_array((C.c_uint8 * NB).from_address(p)).reshape(H, 4 * W)

frame = synth.frame_noise("RGBA", W, H, 1)
bufs = {}
for wc in (0,):
    p, a = host_buf(wc)
    t0 = time.perf_counter(); a[:] = frame; t1 = time.perf_counter()
    bufs[wc] = (p, a)
    print(json.dumps({"what": "cpu fill of the host frame", "write_combined": bool(wc), "ms": round((t1 - t0) * 1e3, 2)}), flush=True)
out_p, out_a = host_buf(0)
d_in = torch.empty((H, 4 * W), dtype=torch.uint8, device="cuda"); d_out = torch.zeros_like(d_in)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
cudart = torch.cuda.cudart()
def memcpy_async(dst, src, kind, stream):
    err = cudart.cudaMemcpyAsync(dst, src, NB, kind, stream.cuda_stream) if hasattr(cudart, "cudaMemcpyAsync") else None
    return err
# torch's cudart binding lacks cudaMemcpyAsync: use ctypes on the runtime that is already loaded
import ctypes.util
lib = None
for name in ("libcudart.so.12", "libcudart.so"):
    try:
        lib = C.CDLL(name); break
    except OSError:
        pass
if lib is None:
    import glob
    cands = glob.glob(os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "cuda_runtime", "lib", "libcudart.so*"))
    lib = C.CDLL(cands[0])
lib.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
H2D, D2H = 1, 2
def run(fn, n=30):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n
for wc in (0,):
    p = bufs[wc][0]
    t = run(lambda: lib.cudaMemcpyAsync(d_in.data_ptr(), p, NB, H2D, s1.cuda_stream))
    print(json.dumps({"what": "memcpy H2D alone", "write_combined_src": bool(wc), "GBps": round(NB / t / 1e9, 2)}), flush=True)
t = run(lambda: lib.cudaMemcpyAsync(out_p, d_out.data_ptr(), NB, D2H, s2.cuda_stream))
print(json.dumps({"what": "memcpy D2H alone", "GBps": round(NB / t / 1e9, 2)}), flush=True)
for wc in (0,):
    p = bufs[wc][0]
    def both():
        lib.cudaMemcpyAsync(d_in.data_ptr(), p, NB, H2D, s1.cuda_stream)
        lib.cudaMemcpyAsync(out_p, d_out.data_ptr(), NB, D2H, s2.cuda_stream)
    t = run(both)
    print(json.dumps({"what": "memcpy H2D + D2H concurrently (two streams)", "write_combined_src": bool(wc), "GBps_each_way": round(NB / t / 1e9, 2),
                      "frames_per_s_ceiling": round(1 / t, 1)}), flush=True)
ctx = b200vfx.Context(0)
k, s, v, sc, of = b200vfx.cube_parse(synth.cube_text_3d(33, "mix")); ctx.colorlut_set_lut(k, s, v, sc, of)
for zc in (1, 0):
    ctx.set_option("zero_copy", zc)
    for wc in (0,):
        p = bufs[wc][0]
        for _ in range(4): ctx.colorlut_process("RGBA", W, H, p, 4 * W, out_p, 4 * W)
        t0 = time.perf_counter()
        for _ in range(24): ctx.colorlut_process("RGBA", W, H, p, 4 * W, out_p, 4 * W)
        t = (time.perf_counter() - t0) / 24
        print(json.dumps({"what": "e2e colorlut_process, pinned host in/out", "zero_copy": zc, "write_combined_src": bool(wc), "fps": round(1 / t, 1),
                          "GBps_each_way": round(NB / t / 1e9, 2)}), flush=True)

