"""numpy model of the format conversions (test infrastructure).  Packed <-> packed is a byte permutation (exact by
definition).  RGB <-> I420/A420 restates the SPEC of gst-plugin-rs_b200/csrc/convert.cuh -- GStreamer's converter is not in
the reference tree, so this pins the product to its own specification only (parity with `videoconvert`: unpinned)."""
import numpy as np

# byte offsets (bpp, r, g, b, a or -1)
PACKED = {"RGBx": (4, 0, 1, 2, -1), "RGBA": (4, 0, 1, 2, 3), "xRGB": (4, 1, 2, 3, -1), "ARGB": (4, 1, 2, 3, 0),
          "BGRx": (4, 2, 1, 0, -1), "BGRA": (4, 2, 1, 0, 3), "xBGR": (4, 3, 2, 1, -1), "ABGR": (4, 3, 2, 1, 0),
          "RGB": (3, 0, 1, 2, -1), "BGR": (3, 2, 1, 0, -1)}


def unpack(fmt, w, h, frame):
    bpp, r, g, b, a = PACKED[fmt]
    px = np.asarray(frame)[:h, :w * bpp].reshape(h, w, bpp)
    alpha = px[:, :, a] if a >= 0 else np.full((h, w), 255, np.uint8)
    return px[:, :, r], px[:, :, g], px[:, :, b], alpha, a >= 0


def convert_packed(src_fmt, dst_fmt, w, h, frame, dst_stride=None, fill=0):
    r, g, b, alpha, has_a = unpack(src_fmt, w, h, frame)
    bpp, ro, go, bo, ao = PACKED[dst_fmt]
    stride = dst_stride or w * bpp
    out = np.full((h, stride), fill, np.uint8)
    px = out[:, :w * bpp].reshape(h, w, bpp)
    px[:, :, ro], px[:, :, go], px[:, :, bo] = r, g, b
    if bpp == 4:
        fourth = 6 - ro - go - bo
        px[:, :, fourth] = alpha if (ao >= 0 and has_a) else 255
    return out


def matrix(kind, height):
    hd = kind == 709 or (kind == 0 and height > 576)
    kr, kb = (0.2126, 0.0722) if hd else (0.299, 0.114)
    kg = 1.0 - kr - kb
    sy, sc = 219.0 / 255.0, 224.0 / 255.0
    q = lambda v: int(np.rint(v * 256.0))
    fwd = dict(yr=q(kr * sy), yg=q(kg * sy), yb=q(kb * sy),
               ur=q(-kr / (2 * (1 - kb)) * sc), ug=q(-kg / (2 * (1 - kb)) * sc), ub=q(0.5 * sc),
               vr=q(0.5 * sc), vg=q(-kg / (2 * (1 - kr)) * sc), vb=q(-kb / (2 * (1 - kr)) * sc))
    inv = dict(y=q(1 / sy), rv=q(2 * (1 - kr) / sc), gu=q(-2 * (1 - kb) * kb / kg / sc), gv=q(-2 * (1 - kr) * kr / kg / sc),
               bu=q(2 * (1 - kb) / sc))
    return fwd, inv


def to_planar(src_fmt, w, h, frame, kind=0, with_alpha=False):
    r, g, b, alpha, _ = unpack(src_fmt, w, h, frame)
    m, _ = matrix(kind, h)
    r, g, b = r.astype(np.int64), g.astype(np.int64), b.astype(np.int64)
    c = lambda v: np.clip(v, 0, 255)
    Y = c((m["yr"] * r + m["yg"] * g + m["yb"] * b + (16 << 8) + 128) >> 8)
    U = c((m["ur"] * r + m["ug"] * g + m["ub"] * b + (128 << 8) + 128) >> 8)
    V = c((m["vr"] * r + m["vg"] * g + m["vb"] * b + (128 << 8) + 128) >> 8)
    ch, cw = (h + 1) // 2, (w + 1) // 2
    def sub(p):
        p = np.pad(p, ((0, 2 * ch - h), (0, 2 * cw - w)), mode="edge")
        return ((p[0::2, 0::2] + p[0::2, 1::2] + p[1::2, 0::2] + p[1::2, 1::2] + 2) >> 2).astype(np.uint8)
    planes = [Y.astype(np.uint8), sub(U), sub(V)]
    if with_alpha:
        planes.append(alpha.copy())
    return planes


def from_planar(planes, dst_fmt, w, h, kind=0):
    _, m = matrix(kind, h)
    Y = planes[0][:h, :w].astype(np.int64) - 16
    U = np.repeat(np.repeat(planes[1], 2, axis=0), 2, axis=1)[:h, :w].astype(np.int64) - 128
    V = np.repeat(np.repeat(planes[2], 2, axis=0), 2, axis=1)[:h, :w].astype(np.int64) - 128
    c = lambda v: np.clip(v, 0, 255).astype(np.uint8)
    r = c((m["y"] * Y + m["rv"] * V + 128) >> 8)
    g = c((m["y"] * Y + m["gu"] * U + m["gv"] * V + 128) >> 8)
    b = c((m["y"] * Y + m["bu"] * U + 128) >> 8)
    bpp, ro, go, bo, ao = PACKED[dst_fmt]
    out = np.zeros((h, w * bpp), np.uint8)
    px = out.reshape(h, w, bpp)
    px[:, :, ro], px[:, :, go], px[:, :, bo] = r, g, b
    if bpp == 4:
        px[:, :, 6 - ro - go - bo] = planes[3][:h, :w] if (ao >= 0 and len(planes) > 3) else 255
    return out
