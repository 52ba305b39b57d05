"""Independent numpy-float32 restatement of SURVEY Appendix A (second opinion for the C oracle).

Written from the arithmetic spec, vectorised; every operator is a separate numpy ufunc on
float32 arrays, i.e. one IEEE rounding per operator, never fused.  Used only by tests.
"""
from __future__ import annotations

import numpy as np

F = np.float32


def _clamp(x, lo, hi):  # f32::clamp, NaN-preserving
    x = np.where(x < F(lo), F(lo), x)
    x = np.where(x > F(hi), F(hi), x)
    return x.astype(F)


def _round_half_away(q):  # f32::round for q >= 0 (and NaN)
    t = np.trunc(q)
    frac = q - t  # exact
    return np.where(frac >= F(0.5), t + F(1), t).astype(F)


def _sat(v, hi):
    v = np.where(np.isnan(v), F(0), v)
    return np.clip(v, 0, hi)


def _norm(c, denom, scale, offset):
    v = c.astype(F) / F(denom)
    n = (v * F(scale)).astype(F)
    n = (n + F(offset)).astype(F)
    return _clamp(n, 0, 1)


def _idx(pos, size):
    fl = np.floor(pos)
    fl = np.where(np.isnan(fl), F(0), fl)
    i0 = np.minimum(np.clip(fl, 0, None).astype(np.int64), size - 1)
    i1 = np.minimum(i0 + 1, size - 1)
    t = (pos - i0.astype(F)).astype(F)
    return i0, i1, t


def _lerp(a, b, t):
    d = (b - a).astype(F)
    p = (d * t).astype(F)
    return (a + p).astype(F)


def colorlut_3d(values, size, scale, offset, r, g, b, denom=255, outmax=255):
    """values (size^3,3) f32; r,g,b integer arrays -> (3, ...) integer outputs."""
    sm1 = F(F(size) - F(1))
    x = (_norm(r, denom, scale[0], offset[0]) * sm1).astype(F)
    y = (_norm(g, denom, scale[1], offset[1]) * sm1).astype(F)
    z = (_norm(b, denom, scale[2], offset[2]) * sm1).astype(F)
    x0, x1, tx = _idx(x, size)
    y0, y1, ty = _idx(y, size)
    z0, z1, tz = _idx(z, size)
    n = size
    out = []
    for k in range(3):
        tab = np.ascontiguousarray(values[:, k])
        at = lambda X, Y, Z: tab[X + Y * n + Z * n * n]
        c00 = _lerp(at(x0, y0, z0), at(x1, y0, z0), tx)
        c10 = _lerp(at(x0, y1, z0), at(x1, y1, z0), tx)
        c01 = _lerp(at(x0, y0, z1), at(x1, y0, z1), tx)
        c11 = _lerp(at(x0, y1, z1), at(x1, y1, z1), tx)
        c0 = _lerp(c00, c10, ty)
        c1 = _lerp(c01, c11, ty)
        o = _lerp(c0, c1, tz)
        q = (_clamp(o, 0, 1) * F(outmax)).astype(F)
        out.append(_sat(_round_half_away(q), outmax).astype(np.int64))
    return out


def colorlut_1d(values, size, scale, offset, c, ch, denom=255, outmax=255):
    sm1 = F(F(size) - F(1))
    x = (_norm(c, denom, scale[ch], offset[ch]) * sm1).astype(F)
    i0, i1, t = _idx(x, size)
    tab = np.ascontiguousarray(values[:, ch])
    o = _lerp(tab[i0], tab[i1], t)
    q = (_clamp(o, 0, 1) * F(outmax)).astype(F)
    return _sat(_round_half_away(q), outmax).astype(np.int64)


def rgb_to_hsv(r8, g8, b8):
    r = r8.astype(F) / F(255)
    g = g8.astype(F) / F(255)
    b = b8.astype(F) / F(255)
    mx = np.maximum(np.maximum(r8, g8), b8)
    mn = np.minimum(np.minimum(r8, g8), b8)
    value = mx.astype(F) / F(255)
    chroma = (value - (mn.astype(F) / F(255))).astype(F)
    eps = F(0.00001)
    with np.errstate(divide="ignore", invalid="ignore"):
        hr = (F(60) * ((g - b).astype(F) / chroma).astype(F)).astype(F)
        hg = (F(60) * (F(2) + ((b - r).astype(F) / chroma).astype(F)).astype(F)).astype(F)
        hb = (F(60) * (F(4) + ((r - g).astype(F) / chroma).astype(F)).astype(F)).astype(F)
    hue = np.where(chroma == 0, F(0),
                   np.where(np.abs(value - r) < eps, hr,
                            np.where(np.abs(value - g) < eps, hg,
                                     np.where(np.abs(value - b) < eps, hb, F(0))))).astype(F)
    hue = np.where(hue < 0, (hue + F(360)).astype(F), hue).astype(F)
    with np.errstate(divide="ignore", invalid="ignore"):
        sat = np.where(value == 0, F(0), (chroma / value).astype(F)).astype(F)
    return np.fmod(hue, F(360)).astype(F), _clamp(sat, 0, 1), _clamp(value, 0, 1)


def hsv_to_rgb(h, s, v):
    c = (v * s).astype(F)
    hp = (h / F(60)).astype(F)
    x = (c * (F(1) - np.abs((np.fmod(hp, F(2)) - F(1)).astype(F))).astype(F)).astype(F)
    z = np.zeros_like(c)
    conds = [hp < 0, hp <= 1, hp <= 2, hp <= 3, hp <= 4, hp <= 5, hp <= 6]
    p0 = np.select(conds, [z, c, x, z, z, x, c], z)
    p1 = np.select(conds, [z, x, c, c, x, z, z], z)
    p2 = np.select(conds, [z, z, z, x, c, c, x], z)
    m = (v - c).astype(F)
    out = []
    for p in (p0, p1, p2):
        q = _clamp(((p + m).astype(F) * F(255)).astype(F), 0, 255)
        out.append(np.trunc(_sat(q, 255)).astype(np.int64))
    return out


def _maxmin_clamp(x):  # hsvutils::Clamp trait: max then min, NaN -> 0
    return np.fmin(np.fmax(x, F(0)), F(1)).astype(F)


def hsvfilter_rgb(r8, g8, b8, hue_shift=0.0, sat_mul=1.0, sat_off=0.0, val_mul=1.0, val_off=0.0):
    h, s, v = rgb_to_hsv(r8, g8, b8)
    h = np.fmod((h + F(hue_shift)).astype(F), F(360)).astype(F)
    h = np.where(h < 0, (h + F(360)).astype(F), h).astype(F)
    s = _maxmin_clamp(((F(sat_mul) * s).astype(F) + F(sat_off)).astype(F))
    v = _maxmin_clamp(((F(val_mul) * v).astype(F) + F(val_off)).astype(F))
    return hsv_to_rgb(h, s, v)


def hsvdetect_rgb(r8, g8, b8, hue_ref=0.0, hue_var=10.0, sat_ref=0.0, sat_var=0.15, val_ref=0.0, val_var=0.3):
    h, s, v = rgb_to_hsv(r8, g8, b8)
    sh = (h + (F(180) - F(hue_ref))).astype(F)
    sh = np.where(sh < 0, (sh + F(360)).astype(F), sh).astype(F)
    sh = np.fmod(sh, F(360)).astype(F)
    return ((np.abs((sh - F(180)).astype(F)) <= F(hue_var))
            & (np.abs((s - F(sat_ref)).astype(F)) <= F(sat_var))
            & (np.abs((v - F(val_ref)).astype(F)) <= F(val_var)))
