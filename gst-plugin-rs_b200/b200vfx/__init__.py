"""b200vfx -- thin ctypes binding of libb200vfx.so (include/b200vfx.h).

This is plumbing for tests/ and bench.py: every call goes straight through the C ABI that a
gstreamer-rs element would bind.  There is no Python or CPU implementation of
the pixel path here -- if the CUDA library is missing or no GPU is present the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
PKG_ROOT = os.path.dirname(HERE)
REPO_ROOT = os.path.dirname(PKG_ROOT)
LIB_PATH = os.environ.get("B200VFX_LIB") or os.path.join(PKG_ROOT, "lib", "libb200vfx.so")   # override: A/B builds
HEADER = os.path.join(REPO_ROOT, "include", "b200vfx.h")

FMT = {"RGBx": 0, "xRGB": 1, "BGRx": 2, "xBGR": 3, "RGBA": 4, "ARGB": 5, "BGRA": 6, "ABGR": 7,
       "RGB": 8, "BGR": 9, "RGBA64_LE": 10, "RGBA64_BE": 11, "I420": 12, "A420": 13}

HASH_ALGO = {"mean": 0, "gradient": 1, "vertgradient": 2, "doublegradient": 3, "blockhash": 4}

OK, ERR_INVALID, ERR_CUDA, ERR_NOT_NEGOTIATED, ERR_UNSUPPORTED, ERR_PARSE, ERR_IO = 0, -1, -2, -3, -4, -5, -6

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-fmad=false", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]
SOURCES = ["csrc/b200vfx.cu", "csrc/cube_parser.cpp", "csrc/elements.cpp", "csrc/colordetect_host.cpp", "csrc/hash_host.cpp"]


class B200VfxError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__("b200vfx error %d: %s" % (code, msg))
        self.code = code
        self.msg = msg


def build(force: bool = False) -> str:
    """Compile libb200vfx.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
    srcs = [os.path.join(PKG_ROOT, s) for s in SOURCES if os.path.exists(os.path.join(PKG_ROOT, s))]
    deps = srcs + [os.path.join(PKG_ROOT, "csrc", f) for f in os.listdir(os.path.join(PKG_ROOT, "csrc"))
                   if f.endswith((".cuh", ".h", ".hpp"))] + [HEADER]
    if not force and os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(d) for d in deps):
        return LIB_PATH
    nvcc = os.environ.get("NVCC") or ("/usr/local/cuda/bin/nvcc" if os.path.exists("/usr/local/cuda/bin/nvcc") else "nvcc")
    os.makedirs(os.path.dirname(LIB_PATH), exist_ok=True)
    cmd = [nvcc] + NVCC_FLAGS + ["-o", LIB_PATH] + srcs
    subprocess.check_call(cmd, cwd=PKG_ROOT)
    return LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        build()
    L = C.CDLL(LIB_PATH)
    vp, ci, cf, cu = C.c_void_p, C.c_int, C.c_float, C.c_uint
    f32p = C.POINTER(C.c_float)
    sigs = {
        "b200vfx_abi_version": ([], ci),
        "b200vfx_device_count": ([], ci),
        "b200vfx_ctx_create": ([C.POINTER(vp), ci], ci),
        "b200vfx_ctx_destroy": ([vp], None),
        "b200vfx_last_error": ([vp], C.c_char_p),
        "b200vfx_ctx_set_stream": ([vp, vp], ci),
        "b200vfx_ctx_synchronize": ([vp], ci),
        "b200vfx_ctx_set_chunk_rows": ([vp, ci], ci),
        "b200vfx_ctx_set_host_async": ([vp, ci], ci),
        "b200vfx_fence_create": ([vp, C.POINTER(vp)], ci),
        "b200vfx_fence_wait": ([vp], ci),
        "b200vfx_fence_query": ([vp], ci),
        "b200vfx_fence_destroy": ([vp], None),
        "b200vfx_ctx_kernel_launches": ([vp], C.c_uint64),
        "b200vfx_ctx_set_option": ([vp, C.c_char_p, ci], ci),
        "b200vfx_host_alloc": ([C.c_size_t], vp),
        "b200vfx_host_free": ([vp], None),
        "b200vfx_device_alloc": ([vp, C.c_size_t], vp),
        "b200vfx_device_free": ([vp, vp], None),
        "b200vfx_upload": ([vp, vp, ci, vp, ci, C.c_size_t, ci], ci),
        "b200vfx_download": ([vp, vp, ci, vp, ci, C.c_size_t, ci], ci),
        "b200vfx_copy_plane": ([vp, vp, ci, vp, ci, C.c_size_t, ci], ci),
        "b200vfx_pointer_is_device": ([vp], ci),
        "b200vfx_debug_pdl_admit": ([vp, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t, ci, C.c_longlong, ci], ci),
        "b200vfx_debug_pdl_reset": ([vp], None),
        "b200vfx_cube_parse": ([C.c_char_p, C.c_size_t, C.POINTER(ci), C.POINTER(ci), C.POINTER(f32p), f32p, f32p,
                                C.c_char_p, C.c_size_t], ci),
        "b200vfx_cube_parse_file": ([C.c_char_p, C.POINTER(ci), C.POINTER(ci), C.POINTER(f32p), f32p, f32p,
                                     C.c_char_p, C.c_size_t], ci),
        "b200vfx_cube_free": ([f32p], None),
        "b200vfx_colorlut_load_file": ([vp, C.c_char_p], ci),
        "b200vfx_colorlut_set_lut": ([vp, ci, ci, f32p, f32p, f32p], ci),
        "b200vfx_colorlut_clear": ([vp], ci),
        "b200vfx_colorlut_set_mode": ([vp, ci], ci),
        "b200vfx_colorlut_process": ([vp, ci, ci, ci, vp, ci, vp, ci], ci),
        "b200vfx_colorlut_process_fmt": ([vp, ci, ci, ci, ci, vp, ci, vp, ci], ci),
        "b200vfx_convert_packed": ([vp, ci, ci, ci, ci, vp, ci, vp, ci], ci),
        "b200vfx_convert_to_planar": ([vp, ci, ci, ci, ci, vp, ci, C.POINTER(vp), C.POINTER(ci), ci], ci),
        "b200vfx_convert_from_planar": ([vp, ci, ci, ci, ci, C.POINTER(vp), C.POINTER(ci), vp, ci, ci], ci),
        "b200vfx_a420_append": ([vp, ci, ci, C.POINTER(vp), C.POINTER(ci), vp, ci, C.POINTER(vp), C.POINTER(ci)], ci),
        "b200vfx_colorlut_process_planar": ([vp, ci, ci, ci, C.POINTER(vp), C.POINTER(ci), C.POINTER(vp), C.POINTER(ci), ci], ci),
        "b200vfx_hsvfilter_process": ([vp, ci, ci, ci, vp, ci] + [cf] * 5, ci),
        "b200vfx_hsvdetector_process": ([vp, ci, ci, ci, ci, vp, ci, vp, ci] + [cf] * 6, ci),
        "b200vfx_roundmask_generate": ([vp, ci, ci, ci, cu, vp], ci),
        "b200vfx_blockhash_sums": ([vp, ci, ci, ci, vp, ci, ci, ci, vp], ci),
        "b200vfx_blockhash_sums_batch": ([vp, ci, ci, ci, ci, C.POINTER(vp), C.POINTER(ci), ci, ci, vp], ci),
        "b200vfx_blockhash_bits": ([vp, ci, ci, ci, ci, vp], None),
        "b200vfx_hash_distance": ([vp, vp, ci], ci),
        "b200vfx_hash_image": ([vp, ci, ci, ci, ci, vp, ci, vp, C.POINTER(ci)], ci),
        "b200vfx_blockhash_sums_f32": ([vp, ci, ci, ci, vp, ci, ci, ci, vp], ci),
        "b200vfx_blockhash_bits_f32": ([vp, ci, ci, ci, ci, vp], None),
        "b200vfx_luma_resize": ([vp, ci, ci, ci, vp, ci, ci, ci, vp], ci),
        "b200vfx_hash_resize_dims": ([ci, C.POINTER(ci), C.POINTER(ci)], ci),
        "b200vfx_debug_resize_taps": ([ci, ci, ci, C.POINTER(ci), f32p, ci], ci),
        "b200vfx_hash_bits_from_luma": ([ci, vp, ci, ci, vp], ci),
        "b200vfx_colordetect_histogram": ([vp, ci, ci, ci, vp, ci, ci, vp], ci),
        "b200vfx_colordetect_palette": ([vp, ci, vp, ci, C.POINTER(ci)], ci),
        "b200vfx_css_color_similar": ([cu, cu, cu], C.c_char_p),
        "b200vfx_peer_alloc": ([vp, C.c_size_t, C.POINTER(vp), C.c_char_p], ci),
        "b200vfx_peer_free": ([vp, vp], ci),
        "b200vfx_peer_open": ([vp, C.c_char_p, C.POINTER(vp)], ci),
        "b200vfx_peer_close": ([vp, vp], ci),
        "b200vfx_peer_enable_access": ([vp, ci], ci),
        "b200vfx_peer_status": ([vp, vp, C.POINTER(C.c_uint32)], ci),
        "b200vfx_colorlut_process_tile_gather": ([vp, ci, ci, ci, vp, ci, ci, ci, C.POINTER(vp), ci, ci, C.POINTER(vp),
                                                  C.c_uint32], ci),
        "b200vfx_colorlut_process_tile_gather_mc": ([vp, ci, ci, ci, vp, ci, ci, ci, C.POINTER(vp), vp, ci, ci,
                                                     C.POINTER(vp), C.c_uint32], ci),
    }
    for name, (args, res) in sigs.items():
        fn = getattr(L, name)
        fn.argtypes = args
        fn.restype = res
    _lib = L
    return L


def exported_symbols_in_header():
    """Names of every function include/b200vfx.h declares (used by the symbol-export test)."""
    import re
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200vfx_[a-z0-9_]+)\s*\(", text)))


def device_count() -> int:
    return lib().b200vfx_device_count()


def _ptr(x) -> int:
    """raw address of a numpy array / torch tensor / int / ctypes pointer"""
    if x is None:
        return 0
    if isinstance(x, int):
        return x
    if hasattr(x, "data_ptr"):
        return int(x.data_ptr())
    if hasattr(x, "ctypes"):
        return int(x.ctypes.data)
    return int(C.cast(x, C.c_void_p).value or 0)


def cube_parse(text):
    """-> (kind, size, values ndarray (n,3) f32, scale (3,), offset (3,)) via the product's C++ parser."""
    import numpy as np
    data = text.encode("utf-8") if isinstance(text, str) else bytes(text)
    kind, size = C.c_int(), C.c_int()
    vals = C.POINTER(C.c_float)()
    scale = np.zeros(3, np.float32)
    offset = np.zeros(3, np.float32)
    err = C.create_string_buffer(512)
    rc = lib().b200vfx_cube_parse(data, len(data), C.byref(kind), C.byref(size), C.byref(vals),
                                  scale.ctypes.data_as(C.POINTER(C.c_float)),
                                  offset.ctypes.data_as(C.POINTER(C.c_float)), err, 512)
    if rc:
        raise B200VfxError(rc, err.value.decode("utf-8", "replace"))
    n = size.value if kind.value == 1 else size.value ** 3
    values = np.ctypeslib.as_array(vals, shape=(n, 3)).copy()
    lib().b200vfx_cube_free(vals)
    return kind.value, size.value, values, scale, offset


class Fence:
    def __init__(self, handle):
        self._h = handle

    def wait(self):
        if lib().b200vfx_fence_wait(self._h) != 0:
            raise B200VfxError(ERR_CUDA, "fence wait failed")

    def done(self) -> bool:
        r = lib().b200vfx_fence_query(self._h)
        if r < 0:
            raise B200VfxError(ERR_CUDA, "fence query failed")
        return r == 1

    def close(self):
        if self._h:
            lib().b200vfx_fence_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Context:
    """One b200vfx_ctx (= one element instance)."""

    def __init__(self, device: int = -1):
        self._h = C.c_void_p()
        rc = lib().b200vfx_ctx_create(C.byref(self._h), device)
        if rc:
            raise B200VfxError(rc, (lib().b200vfx_last_error(None) or b"").decode())

    def close(self):
        if self._h:
            lib().b200vfx_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _chk(self, rc):
        if rc:
            raise B200VfxError(rc, (lib().b200vfx_last_error(self._h) or b"").decode())

    # context ------------------------------------------------------------------------------
    def set_stream(self, cuda_stream: int):
        self._chk(lib().b200vfx_ctx_set_stream(self._h, cuda_stream))

    def synchronize(self):
        self._chk(lib().b200vfx_ctx_synchronize(self._h))

    def set_host_async(self, enable: bool):
        """asynchronous host-frame mode: per-pixel calls on (pinned) host frames return once enqueued; see fence()"""
        self._chk(lib().b200vfx_ctx_set_host_async(self._h, 1 if enable else 0))

    def fence(self) -> "Fence":
        """marks everything submitted so far from host frames; wait() before touching those frames"""
        h = C.c_void_p()
        self._chk(lib().b200vfx_fence_create(self._h, C.byref(h)))
        return Fence(h)

    def set_chunk_rows(self, rows: int):
        self._chk(lib().b200vfx_ctx_set_chunk_rows(self._h, rows))

    def set_option(self, name: str, value: int):
        self._chk(lib().b200vfx_ctx_set_option(self._h, name.encode(), int(value)))

    @property
    def kernel_launches(self) -> int:
        return int(lib().b200vfx_ctx_kernel_launches(self._h))

    # colorlut -----------------------------------------------------------------------------
    def colorlut_load_file(self, location: str):
        self._chk(lib().b200vfx_colorlut_load_file(self._h, location.encode() if location is not None else None))

    def colorlut_set_lut(self, kind, size, values, scale, offset):
        import numpy as np
        v = np.ascontiguousarray(values, np.float32)
        s = np.ascontiguousarray(scale, np.float32)
        o = np.ascontiguousarray(offset, np.float32)
        p = lambda a: a.ctypes.data_as(C.POINTER(C.c_float))
        self._chk(lib().b200vfx_colorlut_set_lut(self._h, kind, size, p(v), p(s), p(o)))

    def colorlut_clear(self):
        self._chk(lib().b200vfx_colorlut_clear(self._h))

    def colorlut_set_mode(self, mode: int):
        self._chk(lib().b200vfx_colorlut_set_mode(self._h, mode))

    def colorlut_process(self, fmt, width, height, src, sstride, dst, dstride):
        self._chk(lib().b200vfx_colorlut_process(self._h, FMT[fmt], width, height, _ptr(src), sstride, _ptr(dst), dstride))

    # multi-GPU tile gather (include/b200vfx.h) -----------------------------------
    def peer_alloc(self, nbytes: int):
        """(device pointer, 64-byte IPC handle) of zeroed device memory that peers may map"""
        p = C.c_void_p()
        h = C.create_string_buffer(64)
        self._chk(lib().b200vfx_peer_alloc(self._h, nbytes, C.byref(p), h))
        return int(p.value), bytes(h.raw)

    def peer_free(self, ptr: int):
        self._chk(lib().b200vfx_peer_free(self._h, ptr))

    def peer_open(self, handle: bytes) -> int:
        p = C.c_void_p()
        self._chk(lib().b200vfx_peer_open(self._h, handle, C.byref(p)))
        return int(p.value)

    def peer_close(self, ptr: int):
        self._chk(lib().b200vfx_peer_close(self._h, ptr))

    def peer_enable_access(self, peer_device: int):
        self._chk(lib().b200vfx_peer_enable_access(self._h, peer_device))

    def peer_status(self, flags: int) -> int:
        e = C.c_uint32(0)
        self._chk(lib().b200vfx_peer_status(self._h, flags, C.byref(e)))
        return int(e.value)

    def colorlut_process_tile_gather(self, fmt, width, tile_rows, src, sstride, world, rank, frames, frame_stride,
                                     frame_row0, flags, epoch, multicast=0):
        """multicast: device address of an NVSwitch multicast mapping bound to every rank's frame buffer (0 = unicast)"""
        fa = (C.c_void_p * world)(*[int(x) for x in frames])
        ga = (C.c_void_p * world)(*[int(x) for x in flags])
        self._chk(lib().b200vfx_colorlut_process_tile_gather_mc(self._h, FMT[fmt], width, tile_rows, _ptr(src), sstride,
                                                                world, rank, fa, C.c_void_p(int(multicast) or None),
                                                                frame_stride, frame_row0, ga, epoch))

    # hsv ----------------------------------------------------------------------------------
    def colorlut_process_fmt(self, in_fmt, out_fmt, width, height, src, sstride, dst, dstride):
        self._chk(lib().b200vfx_colorlut_process_fmt(self._h, FMT[in_fmt], FMT[out_fmt], width, height, _ptr(src), sstride,
                                                     _ptr(dst), dstride))

    def convert_packed(self, src_fmt, dst_fmt, width, height, src, sstride, dst, dstride):
        self._chk(lib().b200vfx_convert_packed(self._h, FMT[src_fmt], FMT[dst_fmt], width, height, _ptr(src), sstride,
                                               _ptr(dst), dstride))

    @staticmethod
    def _planes(planes, strides):
        n = len(planes)
        return (C.c_void_p * n)(*[_ptr(p) for p in planes]), (C.c_int * n)(*strides)

    def convert_to_planar(self, src_fmt, dst_fmt, width, height, src, sstride, planes, strides, matrix=0):
        pp, ss = self._planes(planes, strides)
        self._chk(lib().b200vfx_convert_to_planar(self._h, FMT[src_fmt], FMT[dst_fmt], width, height, _ptr(src), sstride, pp, ss, matrix))

    def convert_from_planar(self, src_fmt, dst_fmt, width, height, planes, strides, dst, dstride, matrix=0):
        pp, ss = self._planes(planes, strides)
        self._chk(lib().b200vfx_convert_from_planar(self._h, FMT[src_fmt], FMT[dst_fmt], width, height, pp, ss, _ptr(dst), dstride, matrix))

    def colorlut_process_planar(self, fmt, width, height, src_planes, src_strides, dst_planes, dst_strides, matrix=0):
        sp, ss = self._planes(src_planes, src_strides)
        dp, ds = self._planes(dst_planes, dst_strides)
        self._chk(lib().b200vfx_colorlut_process_planar(self._h, FMT[fmt], width, height, sp, ss, dp, ds, matrix))

    def a420_append(self, width, height, i420_planes, i420_strides, a8, a8_stride, out_planes, out_strides):
        ip, istr = self._planes(i420_planes, i420_strides)
        op, ostr = self._planes(out_planes, out_strides)
        self._chk(lib().b200vfx_a420_append(self._h, width, height, ip, istr, _ptr(a8), a8_stride, op, ostr))

    def hsvfilter_process(self, fmt, width, height, data, stride, hue_shift=0.0, saturation_mul=1.0,
                          saturation_off=0.0, value_mul=1.0, value_off=0.0):
        self._chk(lib().b200vfx_hsvfilter_process(self._h, FMT[fmt], width, height, _ptr(data), stride, hue_shift,
                                                  saturation_mul, saturation_off, value_mul, value_off))

    def hsvdetector_process(self, in_fmt, out_fmt, width, height, src, sstride, dst, dstride, hue_ref=0.0,
                            hue_var=10.0, saturation_ref=0.0, saturation_var=0.15, value_ref=0.0, value_var=0.3):
        self._chk(lib().b200vfx_hsvdetector_process(self._h, FMT[in_fmt], FMT[out_fmt], width, height, _ptr(src),
                                                    sstride, _ptr(dst), dstride, hue_ref, hue_var, saturation_ref,
                                                    saturation_var, value_ref, value_var))

    # videofx ------------------------------------------------------------------------------
    def roundmask_generate(self, width, height, stride, radius, a8_out):
        self._chk(lib().b200vfx_roundmask_generate(self._h, width, height, stride, radius, _ptr(a8_out)))

    def blockhash_sums(self, fmt, width, height, src, stride, sums, hw=8, hh=8):
        self._chk(lib().b200vfx_blockhash_sums(self._h, FMT[fmt], width, height, _ptr(src), stride, hw, hh, _ptr(sums)))

    def hash_image(self, algo, fmt, width, height, src, stride):
        """-> numpy uint8 array of 0/1 bits (64, or 40 for doublegradient); algo: name or number (videocompare hash-algo)"""
        import numpy as np
        bits = np.zeros(64, np.uint8)
        n = C.c_int()
        self._chk(lib().b200vfx_hash_image(self._h, HASH_ALGO.get(algo, algo), FMT[fmt], width, height, _ptr(src), stride,
                                           bits.ctypes.data, C.byref(n)))
        return bits[:n.value].copy()

    def blockhash_sums_f32(self, fmt, width, height, src, stride, sums, hw=8, hh=8):
        self._chk(lib().b200vfx_blockhash_sums_f32(self._h, FMT[fmt], width, height, _ptr(src), stride, hw, hh, _ptr(sums)))

    def luma_resize(self, fmt, width, height, src, stride, nw, nh):
        import numpy as np
        out = np.zeros((nh, nw), np.uint8)
        self._chk(lib().b200vfx_luma_resize(self._h, FMT[fmt], width, height, _ptr(src), stride, nw, nh, out.ctypes.data))
        return out

    def device_alloc(self, nbytes) -> int:
        p = lib().b200vfx_device_alloc(self._h, nbytes)
        if not p:
            raise B200VfxError(ERR_CUDA, (lib().b200vfx_last_error(self._h) or b"").decode())
        return int(p)

    def device_free(self, p):
        lib().b200vfx_device_free(self._h, p)

    def upload(self, dev_dst, dst_stride, host_src, src_stride, row_bytes, rows):
        self._chk(lib().b200vfx_upload(self._h, _ptr(dev_dst), dst_stride, _ptr(host_src), src_stride, row_bytes, rows))

    def download(self, host_dst, dst_stride, dev_src, src_stride, row_bytes, rows):
        self._chk(lib().b200vfx_download(self._h, _ptr(host_dst), dst_stride, _ptr(dev_src), src_stride, row_bytes, rows))

    def blockhash_sums_batch(self, fmt, width, height, srcs, strides, sums, hw=8, hh=8):
        n = len(srcs)
        ptrs = (C.c_void_p * n)(*[_ptr(x) for x in srcs])
        st = (C.c_int * n)(*strides)
        self._chk(lib().b200vfx_blockhash_sums_batch(self._h, FMT[fmt], width, height, n, ptrs, st, hw, hh, _ptr(sums)))

    def colordetect_histogram(self, fmt, width, height, src, stride, quality, hist):
        self._chk(lib().b200vfx_colordetect_histogram(self._h, FMT[fmt], width, height, _ptr(src), stride, quality, _ptr(hist)))


def colordetect_palette(hist, max_colors):
    """-> list of (r, g, b), most significant first (host-side median cut of the product library)"""
    import numpy as np
    h = np.ascontiguousarray(hist, np.uint32)
    pal = np.zeros(3 * 600, np.uint8)
    n = C.c_int()
    rc = lib().b200vfx_colordetect_palette(h.ctypes.data, max_colors, pal.ctypes.data, 600, C.byref(n))
    if rc != 0:
        raise B200VfxError(rc, "colordetect_palette: invalid argument")
    return [tuple(int(v) for v in pal[3 * i:3 * i + 3]) for i in range(min(n.value, 600))]


def css_color_similar(r, g, b) -> str:
    return lib().b200vfx_css_color_similar(int(r), int(g), int(b)).decode()


def blockhash_bits(sums, width, height, hw=8, hh=8):
    import numpy as np
    s = np.ascontiguousarray(sums, np.uint32)
    bits = np.zeros(hw * hh, np.uint8)
    lib().b200vfx_blockhash_bits(s.ctypes.data, hw, hh, width, height, bits.ctypes.data)
    return bits


def blockhash_bits_f32(sums, width, height, hw=8, hh=8):
    import numpy as np
    s = np.ascontiguousarray(sums, np.float32)
    bits = np.zeros(hw * hh, np.uint8)
    lib().b200vfx_blockhash_bits_f32(s.ctypes.data, hw, hh, width, height, bits.ctypes.data)
    return bits


def hash_bits_from_luma(algo, luma):
    import numpy as np
    l = np.ascontiguousarray(luma, np.uint8)
    bits = np.zeros(96, np.uint8)
    n = lib().b200vfx_hash_bits_from_luma(HASH_ALGO.get(algo, algo), l.ctypes.data, l.shape[1], l.shape[0], bits.ctypes.data)
    return bits[:n].copy()


def hash_resize_dims(algo):
    nw, nh = C.c_int(), C.c_int()
    rc = lib().b200vfx_hash_resize_dims(HASH_ALGO.get(algo, algo), C.byref(nw), C.byref(nh))
    if rc != 0:
        raise B200VfxError(rc, "blockhash does not resize")
    return nw.value, nh.value


def hash_distance(a, b) -> int:
    import numpy as np
    a = np.ascontiguousarray(a, np.uint8)
    b = np.ascontiguousarray(b, np.uint8)
    return int(lib().b200vfx_hash_distance(a.ctypes.data, b.ctypes.data, a.size))
