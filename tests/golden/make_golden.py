#!/usr/bin/env python
"""Generates tests/golden/golden_r01.json.

The reference (Rust) cannot run in this image, so these fixtures come from the CPU oracle AFTER it was pinned against
the reference's own KATs, SURVEY Appendix C and the independent numpy restatement (tests/test_oracle_cpu.py).  They
freeze that state: tests/test_golden.py checks the oracle (CPU) and the CUDA path (GPU) against them, so a later change
to either side cannot drift silently.  Inputs are the deterministic synthetic frames of b200vfx.synth (fixed seeds);
outputs are stored as SHA-256 digests (plus a few explicit pixels for debugging).

    python tests/golden/make_golden.py        # rewrites golden_r01.json
"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "gst-plugin-rs_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np

import oracle_binding as orc
from b200vfx import synth


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def cases():
    """(name, callable(api) -> ndarray).  `api` is a small namespace with colorlut/hsvfilter/hsvdetector/blockhash/mask so the
    same case list can be evaluated by the oracle here and by the CUDA path in tests/test_golden.py."""
    out = []
    luts = {
        "mix33": synth.cube_text_3d(33, "mix"),
        "mix65": synth.cube_text_3d(65, "mix"),
        "mix17dom": synth.cube_text_3d(17, "mix", domain=((-0.25, 0.0, 0.1), (1.5, 0.75, 0.9))),
        "inv2": synth.cube_text_3d(2, "invert"),
        "gamma1d": synth.cube_text_1d(1024, 2.2, domain=((0.0, -0.5, 0.1), (1.0, 1.5, 0.6))),
    }
    sizes = {"small": (317, 43), "sd": (640, 480), "uhd": (3840, 2160)}
    for lname, text in luts.items():
        for sname, (w, h) in sizes.items():
            if sname == "uhd" and lname not in ("mix33", "mix65"):
                continue
            for cname, mk in (("ramps", lambda f, w, h: synth.frame_ramps(f, w, h)), ("noise", lambda f, w, h: synth.frame_noise(f, w, h, 0x5EED0002))):
                out.append(("colorlut/%s/RGBA/%s/%s" % (lname, sname, cname),
                            lambda api, text=text, w=w, h=h, mk=mk: api.colorlut(text, "RGBA", w, h, mk("RGBA", w, h))))
        for fmt in ("RGBA64_LE", "RGBA64_BE"):
            out.append(("colorlut/%s/%s/small/noise" % (lname, fmt),
                        lambda api, text=text, fmt=fmt: api.colorlut(text, fmt, 317, 43, synth.frame_noise(fmt, 317, 43, 0x5EED0002))))
    hf = dict(hue_shift=90.0)
    hf2 = dict(hue_shift=-270.25, sat_mul=1.7, sat_off=-0.2, val_mul=0.8, val_off=0.15)
    for fmt in ("RGBx", "xRGB", "BGRx", "xBGR", "RGBA", "ARGB", "BGRA", "ABGR", "RGB", "BGR"):
        for kwn, kw in (("shift90", hf), ("mixed", hf2)):
            out.append(("hsvfilter/%s/%s/sd/noise" % (fmt, kwn),
                        lambda api, fmt=fmt, kw=kw: api.hsvfilter(fmt, 640, 480, synth.frame_noise(fmt, 640, 480, 0x5EED0001), kw)))
    out.append(("hsvfilter/RGBA/shift90/sd/ramps", lambda api: api.hsvfilter("RGBA", 640, 480, synth.frame_ramps("RGBA", 640, 480), hf)))
    hd = dict(hue_ref=120.0, hue_var=30.0, sat_ref=0.8, sat_var=0.2, val_ref=0.8, val_var=0.2)
    for ifmt in ("RGBx", "xRGB", "BGRx", "xBGR", "RGB", "BGR"):
        for ofmt in ("RGBA", "ARGB", "BGRA", "ABGR"):
            out.append(("hsvdetector/%s-%s/small/noise" % (ifmt, ofmt),
                        lambda api, ifmt=ifmt, ofmt=ofmt: api.hsvdetector(ifmt, ofmt, 317, 43, synth.frame_noise(ifmt, 317, 43, 0x5EED0003), hd)))
    out.append(("hsvdetector/BGRx-RGBA/hd/noise", lambda api: api.hsvdetector("BGRx", "RGBA", 1920, 1080, synth.frame_noise("BGRx", 1920, 1080, 0x5EED0003), hd)))
    out.append(("hsvdetector/BGRx-RGBA/hd/ramps", lambda api: api.hsvdetector("BGRx", "RGBA", 1920, 1080, synth.frame_ramps("BGRx", 1920, 1080), hd)))
    for fmt, (w, h) in (("RGBA", (3840, 2160)), ("RGB", (1920, 1080)), ("RGBA", (64, 48))):
        out.append(("blockhash/%s/%dx%d/noise" % (fmt, w, h), lambda api, fmt=fmt, w=w, h=h: api.blockhash(fmt, w, h, synth.frame_noise(fmt, w, h, 0x5EED0004))))
    out.append(("blockhash/RGBA/3840x2160/ramps", lambda api: api.blockhash("RGBA", 3840, 2160, synth.frame_ramps("RGBA", 3840, 2160))))
    for (w, h, stride, r) in ((1920, 1080, 1920, 64), (64, 50, 64, 12), (33, 17, 36, 5), (640, 480, 640, 0)):
        out.append(("roundmask/%dx%d/r%d" % (w, h, r), lambda api, w=w, h=h, stride=stride, r=r: api.roundmask(w, h, stride, r)))
    for fmt in ("RGB", "RGBA", "ARGB", "BGR", "BGRA"):
        bpp = 3 if fmt in ("RGB", "BGR") else 4
        for (w, h, q, mc) in ((317, 43, 10, 2), (640, 480, 1, 5)):
            out.append(("colordetect/%s/%dx%d/q%d/noise" % (fmt, w, h, q),
                        lambda api, fmt=fmt, w=w, h=h, q=q, mc=mc, bpp=bpp: api.colordetect(fmt, w, h, synth.frame_noise("RGBA", w, h, 0x5EED0006)[:, :bpp * w].copy(), q, mc)))
    out.append(("colordetect/RGBA/3840x2160/q10/ramps", lambda api: api.colordetect("RGBA", 3840, 2160, synth.frame_ramps("RGBA", 3840, 2160), 10, 2)))
    out.append(("colordetect/RGBA/3840x2160/q1/natural", lambda api: api.colordetect("RGBA", 3840, 2160, synth.frame_natural("RGBA", 3840, 2160, 0x5EED0007, amp=3), 1, 8)))
    return out


def _cd_pack(hist, pal):
    return np.concatenate([np.asarray(hist, np.uint32), np.asarray(pal, np.uint32).reshape(-1)])


class OracleApi:
    def colordetect(self, fmt, w, h, frame, quality, max_colors):
        hist = orc.colordetect_histogram(fmt, w, h, frame, quality)
        return _cd_pack(hist, orc.colordetect_palette(hist, max_colors))

    def colorlut(self, text, fmt, w, h, frame):
        return orc.colorlut_apply(orc.cube_parse(text), fmt, w, h, frame, threads=8)

    def hsvfilter(self, fmt, w, h, frame, kw):
        return orc.hsvfilter(fmt, w, h, frame, threads=8, **kw)

    def hsvdetector(self, ifmt, ofmt, w, h, frame, kw):
        return orc.hsvdetector(ifmt, ofmt, w, h, frame, threads=8, **kw)

    def blockhash(self, fmt, w, h, frame):
        return orc.blockhash_sums(fmt, w, h, frame)

    def roundmask(self, w, h, stride, r):
        return orc.roundmask(w, h, stride, r)


def main():
    api = OracleApi()
    golden = {"_about": "SHA-256 of oracle outputs on b200vfx.synth inputs; see make_golden.py", "cases": {}}
    for name, fn in cases():
        a = fn(api)
        golden["cases"][name] = {"sha256": sha(a), "shape": list(a.shape), "head": np.ascontiguousarray(a).reshape(-1)[:12].tolist()}
    # explicit known-answer vectors for readability
    golden["appendix_c"] = {
        "hsvfilter_default_changed_colours": 11093274, "hsvdetector_default_hits": 719, "hsvdetector_config3_hits": 1415062,
        "colorlut_mix33": {"in": [[95, 130, 194], [217, 207, 235], [15, 163, 33]], "out": [[35, 182, 140], [185, 230, 220], [1, 204, 70]]},
    }
    with open(os.path.join(HERE, "golden_r01.json"), "w") as f:
        json.dump(golden, f, indent=1, sort_keys=True)
    print("wrote %d cases" % len(golden["cases"]))


if __name__ == "__main__":
    main()
