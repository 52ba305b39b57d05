#!/usr/bin/env python
"""Generates tests/golden/golden_r02.json: the paths added in round 2 (all videocompare hash algorithms on sizes that are and
are not multiples of 8, colorlut with packed-format converts fused in, colorlut on I420 / A420 planes, packed and planar
conversions).  Same idea as make_golden.py: digests of what the CPU side (oracle + the numpy model of the conversion spec)
produces on deterministic synthetic inputs, frozen so that neither the CPU models nor the CUDA path can drift silently.
tests/test_golden_r02.py checks the CPU side here and the CUDA path on the GPU box (without consulting the oracle).

    python tests/golden/make_golden_r02.py        # rewrites golden_r02.json
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

import np_convert as npc
import oracle_binding as orc
from b200vfx import synth


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def pack_planes(planes):
    return np.concatenate([np.ascontiguousarray(p).reshape(-1) for p in planes])


def planar_input(w, h, seed, with_a):
    """tightly packed I420 / A420 planes of a converted synthetic RGBA frame (so the colours are plausible video)"""
    rgba = synth.frame_natural("RGBA", w, h, seed, amp=6)
    return npc.to_planar("RGBA", w, h, rgba, 0, with_alpha=with_a)


def cases():
    out = []
    for algo in ("mean", "gradient", "vertgradient", "doublegradient", "blockhash"):
        for (w, h) in ((640, 480), (317, 43), (1366, 768)):
            for cname, mk in (("natural", lambda w, h: synth.frame_natural("RGBA", w, h, 0x5EED0010, amp=4)), ("noise", lambda w, h: synth.frame_noise("RGBA", w, h, 0x5EED0011))):
                out.append(("hash/%s/RGBA/%dx%d/%s" % (algo, w, h, cname), lambda api, algo=algo, w=w, h=h, mk=mk: api.hash_image(algo, "RGBA", w, h, mk(w, h))))
        out.append(("hash/%s/RGB/640x360/noise" % algo, lambda api, algo=algo: api.hash_image(algo, "RGB", 640, 360, synth.frame_noise("RGB", 640, 360, 0x5EED0012))))
    luts = {"mix17": synth.cube_text_3d(17, "mix"), "gamma1d": synth.cube_text_1d(256, 2.2)}
    for lname, text in luts.items():
        for (ifmt, ofmt) in (("BGRx", "RGBA"), ("ARGB", "BGRA"), ("RGBA", "xBGR")):
            out.append(("colorlut_fmt/%s/%s-%s/317x43" % (lname, ifmt, ofmt),
                        lambda api, text=text, ifmt=ifmt, ofmt=ofmt: api.colorlut_fmt(text, ifmt, ofmt, 317, 43, npc.convert_packed("RGBA", ifmt, 317, 43, synth.frame_noise("RGBA", 317, 43, 0x5EED0013)))))
        for fmt, (w, h), kind in (("I420", (640, 480), 0), ("A420", (317, 43), 601), ("I420", (1280, 720), 709), ("A420", (64, 48), 0), ("I420", (33, 17), 0)):
            out.append(("colorlut_planar/%s/%s/%dx%d/m%d" % (lname, fmt, w, h, kind),
                        lambda api, text=text, fmt=fmt, w=w, h=h, kind=kind: api.colorlut_planar(text, fmt, w, h, planar_input(w, h, 0x5EED0014, fmt == "A420"), kind)))
    for (sf, df) in (("RGBA", "BGRx"), ("RGB", "ARGB"), ("xBGR", "BGR"), ("BGRA", "RGBA")):
        out.append(("convert_packed/%s-%s/317x43" % (sf, df), lambda api, sf=sf, df=df: api.convert_packed(sf, df, 317, 43, npc.convert_packed("RGBA", sf, 317, 43, synth.frame_noise("RGBA", 317, 43, 0x5EED0015)))))
    for fmt, (w, h), kind in (("RGBA", (640, 480), 0), ("BGRx", (317, 43), 709), ("RGB", (1280, 720), 0)):
        out.append(("to_planar/%s/%dx%d/m%d" % (fmt, w, h, kind), lambda api, fmt=fmt, w=w, h=h, kind=kind: api.to_planar(fmt, w, h, npc.convert_packed("RGBA", fmt, w, h, synth.frame_noise("RGBA", w, h, 0x5EED0016)), kind)))
    return out


class CpuApi:
    """oracle (reference arithmetic) + numpy model of the conversion spec"""

    def hash_image(self, algo, fmt, w, h, frame):
        return np.asarray(orc.hash_image(algo, fmt, w, h, frame), np.uint8)

    def colorlut_fmt(self, text, ifmt, ofmt, w, h, frame):
        cube = orc.cube_parse(text)
        rgba = npc.convert_packed(ifmt, "RGBA", w, h, frame)
        lut = orc.colorlut_apply(cube, "RGBA", w, h, rgba)
        if not (npc.PACKED[ifmt][4] >= 0 and npc.PACKED[ofmt][4] >= 0):
            lut[:, 3::4] = 255
        return npc.convert_packed("RGBA", ofmt, w, h, lut)

    def colorlut_planar(self, text, fmt, w, h, planes, kind):
        cube = orc.cube_parse(text)
        rgba = npc.from_planar(planes, "RGBA", w, h, kind)
        lut = orc.colorlut_apply(cube, "RGBA", w, h, rgba)
        return pack_planes(npc.to_planar("RGBA", w, h, lut, kind, with_alpha=fmt == "A420"))

    def convert_packed(self, sf, df, w, h, frame):
        return npc.convert_packed(sf, df, w, h, frame)

    def to_planar(self, fmt, w, h, frame, kind):
        return pack_planes(npc.to_planar(fmt, w, h, frame, kind, with_alpha=False))


def main():
    api = CpuApi()
    golden = {"_about": "SHA-256 of CPU-side outputs (oracle + conversion spec) on b200vfx.synth inputs; see make_golden_r02.py", "cases": {}}
    for name, fn in cases():
        a = fn(api)
        golden["cases"][name] = {"sha256": sha(a), "shape": list(np.asarray(a).shape), "head": np.ascontiguousarray(a).reshape(-1)[:12].tolist()}
    with open(os.path.join(HERE, "golden_r02.json"), "w") as f:
        json.dump(golden, f, indent=1, sort_keys=True)
    print("wrote %d cases" % len(golden["cases"]))


if __name__ == "__main__":
    main()
