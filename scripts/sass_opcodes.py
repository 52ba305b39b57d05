#!/usr/bin/env python3
"""Per-kernel SASS opcode histogram of libb200vfx.so (cuobjdump -sass), the evidence behind DESIGN.md's
instruction-count statements: which kernels carry UBLKCP / SYNCS (TMA + mbarrier), FADD2 / FMUL2 (packed f32x2),
LDG.E.ENL2.256 (256-bit table loads), and that no FFMA sits in the colorlut interpolation code.

usage: sass_opcodes.py [--lib PATH] [--match REGEX] [--full] [--out FILE]
"""
import argparse
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEFAULT_LIB = os.path.join(ROOT, "gst-plugin-rs_b200", "lib", "libb200vfx.so")


def kernels(lib):
    """yields (demangled name, [opcode, ...]) per kernel of the sm_100a cubin"""
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    names, bodies = [], []
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            names.append(m.group(1))
            bodies.append([])
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and bodies:
            bodies[-1].append(m.group(1))
    dem = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return list(zip(dem, bodies))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lib", default=DEFAULT_LIB)
    ap.add_argument("--match", default=".")
    ap.add_argument("--full", action="store_true", help="full opcode names (with modifiers) instead of base mnemonics")
    ap.add_argument("--out")
    a = ap.parse_args()
    out = open(a.out, "w") if a.out else sys.stdout
    rx = re.compile(a.match)
    for name, ops in kernels(a.lib):
        if not rx.search(name):
            continue
        short = re.sub(r"\(.*", "", name).replace("void b200vfx::", "")
        hist = collections.Counter(op if a.full else op.split(".")[0] for op in ops)
        flags = [k for k in ("UBLKCP", "SYNCS", "FADD2", "FMUL2", "FFMA2", "FFMA", "UTMALDG", "UTMASTG", "REDUX", "ATOMS", "MUFU")
                 if any(op.split(".")[0] == k for op in ops)]
        if any(op.startswith("LDG.E.ENL2.256") for op in ops):
            flags.append("LDG.E.ENL2.256")
        print("%s  [%d instr]  %s" % (short, len(ops), " ".join(flags)), file=out)
        print("    " + " ".join("%s:%d" % kv for kv in hist.most_common()), file=out)


if __name__ == "__main__":
    main()
