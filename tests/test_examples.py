"""The C ABI really is a C ABI: both headers compile as C99, and a plain-C client (examples/colorlut_c_abi.c) links
against libb200vfx.so and -- on a GPU box -- produces the oracle's frame."""
import os
import shutil
import subprocess

import numpy as np
import pytest

import b200vfx
import oracle_binding as orc
from b200vfx import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GCC = shutil.which("gcc")


def fnv1a(a):
    h = 2166136261
    for chunk in np.array_split(np.ascontiguousarray(a).reshape(-1), max(1, a.size // (1 << 16))):
        for b in chunk.tolist():
            h = ((h ^ b) * 16777619) & 0xFFFFFFFF
    return h


@pytest.mark.skipif(GCC is None, reason="gcc not available")
def test_headers_compile_as_c99(tmp_path):
    src = tmp_path / "hdr.c"
    src.write_text('#include "b200vfx.h"\n#include "b200gst.h"\nint main(void) { return B200VFX_ABI_VERSION + B200GST_FLOW_OK - 1; }\n')
    subprocess.check_call([GCC, "-std=c99", "-pedantic", "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), str(src)])


def build_example(tmp_path):
    b200vfx.build()
    exe = str(tmp_path / "colorlut_c_abi")
    libdir = os.path.dirname(b200vfx.LIB_PATH)
    subprocess.check_call([GCC, "-std=c99", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "colorlut_c_abi.c"),
                           "-o", exe, "-L", libdir, "-lb200vfx", "-Wl,-rpath," + libdir])
    return exe


@pytest.mark.skipif(GCC is None, reason="gcc not available")
def test_c_client_links_and_fails_loudly_without_gpu(tmp_path):
    exe = build_example(tmp_path)
    cube = tmp_path / "m.cube"
    cube.write_text(synth.cube_text_3d(9, "mix"))
    r = subprocess.run([exe, str(cube), "64", "16"], capture_output=True, text=True)
    if b200vfx.device_count() <= 0:
        assert r.returncode == 1 and "no CUDA device" in r.stderr      # no CPU fallback behind the ABI
    else:
        assert r.returncode == 0, r.stderr


@pytest.mark.gpu
@pytest.mark.skipif(GCC is None, reason="gcc not available")
def test_c_client_matches_oracle(tmp_path):
    exe = build_example(tmp_path)
    cube = tmp_path / "m.cube"
    cube.write_text(synth.cube_text_3d(17, "mix"))
    w, h = 640, 360
    r = subprocess.run([exe, str(cube), str(w), str(h)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    frame = synth.pcg32(w * h, 0x5EED0002).astype("<u4").view(np.uint8).reshape(h, 4 * w)
    exp = orc.colorlut_apply(orc.cube_parse(cube.read_text()), "RGBA", w, h, frame)
    fields = dict(f.split("=") for f in r.stdout.split()[1:])
    assert int(fields["in"], 16) == fnv1a(frame), "the C client's frame generator diverged from synth.pcg32"
    assert int(fields["out"], 16) == fnv1a(exp)
    assert int(fields["launches"]) >= 3
