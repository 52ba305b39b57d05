"""Deterministic synthetic inputs -- stand-ins for `videotestsrc`.

The reference's own tests use `videotestsrc pattern=red|snow`
(video/videofx/tests/videocompare.rs:61-70,110); there is no GStreamer here, so
frames are generated directly:

* frame A "ramps"  : spatially coherent (best case for LUT locality)
* frame B "noise"  : PCG32 XSH-RR bytes (worst case, moral equivalent of `snow`)
* frame C "natural": ramps + small per-channel PCG noise (proxy for camera content)
* solid frames     : `pattern=red` etc.

plus `.cube` text generators (identity / invert / "mix" LUT of SURVEY 8(d)#2, 1D gamma).
"""
from __future__ import annotations

import numpy as np

_MULT = np.uint64(6364136223846793005)
_M64 = (1 << 64) - 1


def pcg32(n: int, seed: int) -> np.ndarray:
    """n 32-bit PCG32 (XSH-RR) draws, standard pcg32_srandom(seed, seed) seeding."""
    inc = ((seed << 1) | 1) & _M64
    mult = int(_MULT)
    state = 0
    state = (state * mult + inc) & _M64
    state = (state + seed) & _M64
    state = (state * mult + inc) & _M64
    if n <= 0:
        return np.zeros(0, np.uint32)
    block = min(n, 1 << 14)
    first = np.empty(block, np.uint64)
    s = state
    for i in range(block):
        first[i] = s
        s = (s * mult + inc) & _M64
    # jump-ahead by `block`: s_{k+block} = ab * s_k + cb
    ab, cb = 1, 0
    for _ in range(block):
        cb = (cb * mult + inc) & _M64
        ab = (ab * mult) & _M64
    nblocks = -(-n // block)
    states = np.empty((nblocks, block), np.uint64)
    states[0] = first
    ab_u, cb_u = np.uint64(ab), np.uint64(cb)
    with np.errstate(over="ignore"):
        for k in range(1, nblocks):
            states[k] = states[k - 1] * ab_u + cb_u
        old = states.reshape(-1)[:n]
        xorshifted = (((old >> np.uint64(18)) ^ old) >> np.uint64(27)).astype(np.uint32)
        rot = (old >> np.uint64(59)).astype(np.uint32)
        out = (xorshifted >> rot) | (xorshifted << ((np.uint32(32) - rot) & np.uint32(31)))
    return out.astype(np.uint32)


# canonical R,G,B,A planes -----------------------------------------------------------------
def _ramps_planes(w: int, h: int, maxv: int):
    x = np.arange(w, dtype=np.int64)[None, :]
    y = np.arange(h, dtype=np.int64)[:, None]
    r = (maxv * x) // max(w - 1, 1) + 0 * y
    g = (maxv * y) // max(h - 1, 1) + 0 * x
    b = (maxv * (x + y)) // max(w + h - 2, 1)
    a = maxv - (((x ^ y) & 255) * (maxv // 255))
    return r, g, b, a


_ORDER = {  # format -> canonical channel index per memory byte ('x' padding gets alpha plane)
    "RGBx": "RGBA", "RGBA": "RGBA", "xRGB": "ARGB", "ARGB": "ARGB",
    "BGRx": "BGRA", "BGRA": "BGRA", "xBGR": "ABGR", "ABGR": "ABGR",
    "RGB": "RGB", "BGR": "BGR",
}


def default_stride(fmt: str, w: int) -> int:
    if fmt in ("RGB", "BGR"):
        return (3 * w + 3) & ~3
    if fmt in ("RGBA64_LE", "RGBA64_BE"):
        return 8 * w
    return 4 * w


def _pack8(planes, fmt: str, w: int, h: int, stride: int | None, pad_byte: int = 0xA5) -> np.ndarray:
    order = _ORDER[fmt]
    bpp = len(order)
    stride = stride or default_stride(fmt, w)
    assert stride >= bpp * w
    buf = np.full((h, stride), pad_byte, np.uint8)
    idx = {"R": 0, "G": 1, "B": 2, "A": 3}
    for k, ch in enumerate(order):
        buf[:, k:bpp * w:bpp] = planes[idx[ch]].astype(np.uint8)
    return buf


def frame_ramps(fmt: str, w: int, h: int, stride: int | None = None) -> np.ndarray:
    """Frame A. Returns (h, stride) uint8 array (row padding filled with 0xA5)."""
    if fmt in ("RGBA64_LE", "RGBA64_BE"):
        r, g, b, a = _ramps_planes(w, h, 65535)
        stride = stride or 8 * w
        dt = "<u2" if fmt.endswith("LE") else ">u2"
        buf = np.full((h, stride), 0xA5, np.uint8)
        px = np.stack([r, g, b, a], axis=-1).astype(dt)
        buf[:, :8 * w] = px.reshape(h, -1).view(np.uint8)
        return buf
    return _pack8(_ramps_planes(w, h, 255), fmt, w, h, stride)


def frame_noise(fmt: str, w: int, h: int, seed: int, stride: int | None = None) -> np.ndarray:
    """Frame B: PCG32 bytes, one 32-bit draw per 8-bit pixel (two per 16-bit pixel)."""
    if fmt in ("RGBA64_LE", "RGBA64_BE"):
        stride = stride or 8 * w
        d = pcg32(2 * w * h, seed).astype("<u4")
        buf = np.full((h, stride), 0xA5, np.uint8)
        buf[:, :8 * w] = d.view(np.uint8).reshape(h, 8 * w)
        return buf
    bpp = len(_ORDER[fmt])
    stride = stride or default_stride(fmt, w)
    d = pcg32(w * h, seed).astype("<u4").view(np.uint8).reshape(h, w, 4)
    buf = np.full((h, stride), 0xA5, np.uint8)
    buf[:, :bpp * w] = d[:, :, :bpp].reshape(h, bpp * w)
    return buf


def frame_natural(fmt: str, w: int, h: int, seed: int, amp: int = 3, stride: int | None = None) -> np.ndarray:
    """Frame C: ramps plus uniform noise in [-amp, amp] per colour channel (8-bit formats)."""
    r, g, b, a = _ramps_planes(w, h, 255)
    d = pcg32(w * h, seed).reshape(h, w)
    span = 2 * amp + 1
    nr = (d & 0xFF).astype(np.int64) % span - amp
    ng = ((d >> 8) & 0xFF).astype(np.int64) % span - amp
    nb = ((d >> 16) & 0xFF).astype(np.int64) % span - amp
    planes = (np.clip(r + nr, 0, 255), np.clip(g + ng, 0, 255), np.clip(b + nb, 0, 255), a + 0 * r)
    return _pack8(planes, fmt, w, h, stride)


def frame_solid(fmt: str, w: int, h: int, rgba=(255, 0, 0, 255), stride: int | None = None) -> np.ndarray:
    planes = tuple(np.full((h, w), v, np.int64) for v in rgba)
    return _pack8(planes, fmt, w, h, stride)


# .cube generators ---------------------------------------------------------------------------
def _fmt6(v) -> str:
    return "%.6f" % float(v)


def cube_text_3d(n: int, kind: str = "mix", domain=None, title: str | None = None) -> str:
    """`mix`: entry(x,y,z) = ((x/(n-1))^2, sqrt(y/(n-1)), (x+y+z)/(3(n-1))) printed with 6 decimals
    (SURVEY 8(d) config 2); `identity`; `invert`."""
    lines = []
    if title:
        lines.append('TITLE "%s"' % title)
    lines.append("LUT_3D_SIZE %d" % n)
    if domain is not None:
        lo, hi = domain
        lines.append("DOMAIN_MIN %s %s %s" % tuple(_fmt6(v) for v in lo))
        lines.append("DOMAIN_MAX %s %s %s" % tuple(_fmt6(v) for v in hi))
    m = float(n - 1)
    for z in range(n):
        for y in range(n):
            for x in range(n):
                if kind == "identity":
                    v = (x / m, y / m, z / m)
                elif kind == "invert":
                    v = (1 - x / m, 1 - y / m, 1 - z / m)
                elif kind == "mix":
                    v = ((x / m) ** 2, (y / m) ** 0.5, (x + y + z) / (3 * m))
                else:
                    raise ValueError(kind)
                lines.append("%s %s %s" % tuple(_fmt6(c) for c in v))
    return "\n".join(lines) + "\n"


def cube_text_1d(n: int, gamma: float = 2.0, domain=None) -> str:
    lines = ["LUT_1D_SIZE %d" % n]
    if domain is not None:
        lo, hi = domain
        lines.append("DOMAIN_MIN %s %s %s" % tuple(_fmt6(v) for v in lo))
        lines.append("DOMAIN_MAX %s %s %s" % tuple(_fmt6(v) for v in hi))
    m = float(n - 1)
    for i in range(n):
        t = i / m
        lines.append("%s %s %s" % (_fmt6(t ** gamma), _fmt6(t ** (1 / gamma)), _fmt6(1 - t)))
    return "\n".join(lines) + "\n"


def lut_values_3d(n: int, kind: str = "mix") -> np.ndarray:
    """f32 table built directly in f32 (the SURVEY Appendix C construction), shape (n^3, 3)."""
    m = np.float32(n - 1)
    g = np.arange(n, dtype=np.float32)
    z, y, x = np.meshgrid(g, g, g, indexing="ij")
    if kind == "identity":
        v = (x / m, y / m, z / m)
    elif kind == "invert":
        one = np.float32(1)
        v = (one - x / m, one - y / m, one - z / m)
    elif kind == "mix":
        v = ((x / m) * (x / m), np.sqrt(y / m), (x + y + z) / (np.float32(3) * m))
    else:
        raise ValueError(kind)
    return np.stack([c.astype(np.float32).reshape(-1) for c in v], axis=-1)
