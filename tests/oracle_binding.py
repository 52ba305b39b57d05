"""ctypes binding of oracle/libvfx_oracle.so (the CPU oracle -- test infrastructure only).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "libvfx_oracle.so")

FMT = {"RGBx": 0, "xRGB": 1, "BGRx": 2, "xBGR": 3, "RGBA": 4, "ARGB": 5, "BGRA": 6, "ABGR": 7,
       "RGB": 8, "BGR": 9, "RGBA64_LE": 10, "RGBA64_BE": 11, "I420": 12, "A420": 13}


def build(force: bool = False) -> str:
    srcs = [os.path.join(ORACLE_DIR, f) for f in ("vfx_oracle.c", "vfx_oracle_hash.c", "vfx_oracle.h")]
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < max(os.path.getmtime(f) for f in srcs):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"])
    return LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB_PATH)
        f32p, u8p, u32p = C.POINTER(C.c_float), C.POINTER(C.c_uint8), C.POINTER(C.c_uint32)
        _lib.orc_cube_parse.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                        C.POINTER(f32p), f32p, f32p, C.c_char_p, C.c_size_t]
        _lib.orc_cube_parse_file.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                             C.POINTER(f32p), f32p, f32p, C.c_char_p, C.c_size_t]
        _lib.orc_free.argtypes = [C.c_void_p]
        _lib.orc_colorlut_apply.argtypes = [C.c_int, C.c_int, f32p, f32p, f32p, C.c_int, C.c_int, C.c_int,
                                            C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int]
        _lib.orc_hsvfilter.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int] + [C.c_float] * 5 + [C.c_int]
        _lib.orc_hsvdetector.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                                         C.c_int] + [C.c_float] * 6 + [C.c_int]
        _lib.orc_blockhash_sums.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                            u32p, C.c_int]
        _lib.orc_blockhash_bits.argtypes = [u32p, C.c_int, C.c_int, C.c_int, C.c_int, u8p]
        _lib.orc_hamming.argtypes = [u8p, u8p, C.c_int]
        _lib.orc_luma_resize.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        _lib.orc_hash_resize_dims.argtypes = [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        _lib.orc_hash_resize_dims.restype = None
        _lib.orc_hash_bits_from_luma.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        _lib.orc_blockhash_sums_f32.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        _lib.orc_blockhash_bits_f32.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        _lib.orc_blockhash_bits_f32.restype = None
        _lib.orc_roundmask.argtypes = [C.c_int, C.c_int, C.c_int, C.c_uint, C.c_void_p]
        _lib.orc_colordetect_histogram.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        _lib.orc_colordetect_palette.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        _lib.orc_css_similar.argtypes = [C.c_uint, C.c_uint, C.c_uint]
        _lib.orc_css_similar.restype = C.c_char_p
        for n in ("orc_hsv_from_rgb", "orc_hsv_from_bgr"):
            getattr(_lib, n).argtypes = [u8p, f32p]
            getattr(_lib, n).restype = None
        for n in ("orc_hsv_to_rgb", "orc_hsv_to_bgr"):
            getattr(_lib, n).argtypes = [f32p, u8p]
            getattr(_lib, n).restype = None
    return _lib


def _f32p(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


class CubeError(Exception):
    def __init__(self, code, msg):
        super().__init__(msg)
        self.code = code


class Cube:
    """Parsed LUT: kind (1|3), size, values (n,3) f32, scale/offset (3,) f32."""

    def __init__(self, kind, size, values, scale, offset):
        self.kind, self.size = kind, size
        self.values = np.ascontiguousarray(values, np.float32)
        self.scale = np.ascontiguousarray(scale, np.float32)
        self.offset = np.ascontiguousarray(offset, np.float32)


def cube_parse(text) -> Cube:
    data = text.encode("utf-8") if isinstance(text, str) else bytes(text)
    kind, size = C.c_int(), C.c_int()
    vals = C.POINTER(C.c_float)()
    scale = np.zeros(3, np.float32)
    offset = np.zeros(3, np.float32)
    err = C.create_string_buffer(512)
    rc = lib().orc_cube_parse(data, len(data), C.byref(kind), C.byref(size), C.byref(vals), _f32p(scale),
                              _f32p(offset), err, 512)
    if rc != 0:
        raise CubeError(rc, err.value.decode("utf-8", "replace"))
    n = size.value if kind.value == 1 else size.value ** 3
    values = np.ctypeslib.as_array(vals, shape=(n, 3)).copy()
    lib().orc_free(vals)
    return Cube(kind.value, size.value, values, scale, offset)


def cube_from_values(kind, size, values, scale=(1, 1, 1), offset=(-0.0, -0.0, -0.0)) -> Cube:
    return Cube(kind, size, np.asarray(values, np.float32).reshape(-1, 3), np.asarray(scale, np.float32),
                np.asarray(offset, np.float32))


def colorlut_apply(cube: Cube, fmt: str, width: int, height: int, src: np.ndarray, dst_stride=None,
                   threads: int = 1, dst_fill: int = 0x5A, out: np.ndarray | None = None) -> np.ndarray:
    src = np.ascontiguousarray(src, np.uint8)
    sstride = src.shape[1]
    dstride = dst_stride or sstride
    dst = out if out is not None else np.full((height, dstride), dst_fill, np.uint8)
    rc = lib().orc_colorlut_apply(cube.kind, cube.size, _f32p(cube.values), _f32p(cube.scale), _f32p(cube.offset),
                                  FMT[fmt], width, height, src.ctypes.data, sstride, dst.ctypes.data, dstride,
                                  threads)
    if rc:
        raise RuntimeError("orc_colorlut_apply rc=%d" % rc)
    return dst


def hsvfilter(fmt, width, height, data: np.ndarray, hue_shift=0.0, sat_mul=1.0, sat_off=0.0, val_mul=1.0,
              val_off=0.0, threads=1) -> np.ndarray:
    out = np.ascontiguousarray(data, np.uint8).copy()
    rc = lib().orc_hsvfilter(FMT[fmt], width, height, out.ctypes.data, out.shape[1], hue_shift, sat_mul, sat_off,
                             val_mul, val_off, threads)
    if rc:
        raise RuntimeError("orc_hsvfilter rc=%d" % rc)
    return out


def hsvdetector(in_fmt, out_fmt, width, height, src: np.ndarray, dst_stride=None, hue_ref=0.0, hue_var=10.0,
                sat_ref=0.0, sat_var=0.15, val_ref=0.0, val_var=0.3, threads=1, dst_fill=0x5A) -> np.ndarray:
    src = np.ascontiguousarray(src, np.uint8)
    dstride = dst_stride or 4 * width
    dst = np.full((height, dstride), dst_fill, np.uint8)
    rc = lib().orc_hsvdetector(FMT[in_fmt], FMT[out_fmt], width, height, src.ctypes.data, src.shape[1],
                               dst.ctypes.data, dstride, hue_ref, hue_var, sat_ref, sat_var, val_ref, val_var,
                               threads)
    if rc:
        raise RuntimeError("orc_hsvdetector rc=%d" % rc)
    return dst


def blockhash_sums(fmt, width, height, src: np.ndarray, hw=8, hh=8) -> np.ndarray:
    src = np.ascontiguousarray(src, np.uint8)
    sums = np.zeros(hw * hh, np.uint32)
    rc = lib().orc_blockhash_sums(FMT[fmt], width, height, src.ctypes.data, src.shape[1], hw, hh,
                                  sums.ctypes.data_as(C.POINTER(C.c_uint32)), 1)
    if rc:
        raise RuntimeError("orc_blockhash_sums rc=%d" % rc)
    return sums


def blockhash_bits(sums: np.ndarray, width, height, hw=8, hh=8) -> np.ndarray:
    sums = np.ascontiguousarray(sums, np.uint32)
    bits = np.zeros(hw * hh, np.uint8)
    lib().orc_blockhash_bits(sums.ctypes.data_as(C.POINTER(C.c_uint32)), hw, hh, width, height,
                             bits.ctypes.data_as(C.POINTER(C.c_uint8)))
    return bits


def roundmask(width, height, stride, radius) -> np.ndarray:
    rows = (height + 1) & ~1
    a8 = np.full((rows, stride), 0x77, np.uint8)
    rc = lib().orc_roundmask(width, height, stride, radius, a8.ctypes.data)
    if rc:
        raise RuntimeError("orc_roundmask rc=%d" % rc)
    return a8


def hsv_from(rgb, bgr=False):
    p = (C.c_uint8 * 3)(*rgb)
    o = (C.c_float * 3)()
    (lib().orc_hsv_from_bgr if bgr else lib().orc_hsv_from_rgb)(p, o)
    return [o[0], o[1], o[2]]


def hsv_to(hsv, bgr=False):
    p = (C.c_float * 3)(*hsv)
    o = (C.c_uint8 * 3)()
    (lib().orc_hsv_to_bgr if bgr else lib().orc_hsv_to_rgb)(p, o)
    return [o[0], o[1], o[2]]


def colordetect_histogram(fmt, width, height, src: np.ndarray, quality=10) -> np.ndarray:
    """src: (height, stride) uint8 plane; the reference samples the flat stride*height slice"""
    src = np.ascontiguousarray(src, np.uint8)
    hist = np.zeros(32768, np.uint32)
    stride = src.shape[1] if src.ndim == 2 else 0
    rc = lib().orc_colordetect_histogram(FMT[fmt], width, height, src.ctypes.data, stride, quality, hist.ctypes.data)
    if rc:
        raise RuntimeError("orc_colordetect_histogram rc=%d" % rc)
    return hist


def colordetect_palette(hist: np.ndarray, max_colors=2):
    hist = np.ascontiguousarray(hist, np.uint32)
    pal = np.zeros(3 * 600, np.uint8)
    n = lib().orc_colordetect_palette(hist.ctypes.data, max_colors, pal.ctypes.data, 600)
    if n < 0:
        raise RuntimeError("orc_colordetect_palette rc=%d" % n)
    return [tuple(int(v) for v in pal[3 * i:3 * i + 3]) for i in range(min(n, 600))]


def css_similar(r, g, b) -> str:
    return lib().orc_css_similar(int(r), int(g), int(b)).decode()


HASH_ALGO = {"mean": 0, "gradient": 1, "vertgradient": 2, "doublegradient": 3, "blockhash": 4}


def luma_resize(fmt, width, height, frame, nw, nh) -> np.ndarray:
    """image::imageops::grayscale + resize(Lanczos3) as recalled (vfx_oracle_hash.c): nh x nw luma bytes"""
    f = np.ascontiguousarray(frame)
    out = np.zeros((nh, nw), np.uint8)
    rc = lib().orc_luma_resize(FMT[fmt], width, height, f.ctypes.data, f.shape[1], nw, nh, out.ctypes.data)
    assert rc == 0, rc
    return out


def hash_image(algo, fmt, width, height, frame) -> np.ndarray:
    """HasherEngine::hash_image for any of the five algorithms and any frame size: array of 0/1 bits"""
    a = HASH_ALGO.get(algo, algo)
    f = np.ascontiguousarray(frame)
    if a == 4:
        if width % 8 == 0 and height % 8 == 0:
            return blockhash_bits(blockhash_sums(fmt, width, height, f), width, height)
        sums = np.zeros(64, np.float32)
        rc = lib().orc_blockhash_sums_f32(FMT[fmt], width, height, f.ctypes.data, f.shape[1], 8, 8, sums.ctypes.data)
        assert rc == 0, rc
        bits = np.zeros(64, np.uint8)
        lib().orc_blockhash_bits_f32(sums.ctypes.data, 8, 8, width, height, bits.ctypes.data)
        return bits
    nw, nh = C.c_int(), C.c_int()
    lib().orc_hash_resize_dims(a, C.byref(nw), C.byref(nh))
    luma = luma_resize(fmt, width, height, f, nw.value, nh.value)
    bits = np.zeros(96, np.uint8)
    n = lib().orc_hash_bits_from_luma(a, luma.ctypes.data, nw.value, nh.value, bits.ctypes.data)
    return bits[:n].copy()


def blockhash_sums_f32(fmt, width, height, frame) -> np.ndarray:
    f = np.ascontiguousarray(frame)
    sums = np.zeros(64, np.float32)
    rc = lib().orc_blockhash_sums_f32(FMT[fmt], width, height, f.ctypes.data, f.shape[1], 8, 8, sums.ctypes.data)
    assert rc == 0, rc
    return sums
