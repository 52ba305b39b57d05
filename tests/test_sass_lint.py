"""SASS lint (CPU, needs only cuobjdump): the reference never fuses `a + (b - a) * t`, so the interpolation code of
the colorlut kernels must not contain FFMA; the only FMAs allowed are the two Newton-correction FMAs of the exact
x/65535 division in the RGBA64 kernels (scalar for B, one packed f32x2 pair for R,G, x {ident, general domain}) and whatever the compiler's
own IEEE division / fmodf sequences use in the hsv kernels.  Also proves the Blackwell-native pieces are really in
the binary: UBLKCP (cp.async.bulk through the TMA engine), SYNCS (mbarrier) and 256-bit LDG."""
import re
import shutil
import subprocess

import pytest

import b200vfx

cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"


@pytest.fixture(scope="module")
def sass():
    try:
        out = subprocess.run([cuobjdump, "-sass", b200vfx.LIB_PATH], capture_output=True, text=True, check=True).stdout
    except Exception as e:  # pragma: no cover
        pytest.skip("cuobjdump unavailable: %s" % e)
    funcs, cur = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = []
        elif cur and re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
            funcs[cur].append(line.split("*/", 1)[1].strip())
    return funcs


def count(funcs, name_part, op):
    n = 0
    for name, body in funcs.items():
        if name_part in name:
            n += sum(1 for ins in body if re.match(r"(@!?U?P\d+\s+)?" + op + r"\b", ins))
    return n


def count_fused_ffma2(funcs, name_part):
    """FFMA2 instructions whose addend is NOT the opaque -0.0 kernel parameter (a uniform register): real fusions.
    mul2_exact (colorlut_math.cuh) writes every packed product as FFMA2 d, a, b, UR<n>.F32 with UR<n> = -0.0, which is
    the correctly rounded product and cannot be merged with the FADD2 that consumes it."""
    n = 0
    for name, body in funcs.items():
        if name_part in name:
            for ins in body:
                if re.match(r"(@!?U?P\d+\s+)?FFMA2\b", ins) and not re.search(r",\s*-?UR\d+\.F32\s*;", ins):
                    n += 1
    return n


def test_no_fused_multiply_add_in_u8_colorlut_paths(sass):
    for k in ("colorlut_memo_build_kernel", "colorlut_direct_kernelILi0E", "colorlut_memo1d_build_kernel"):
        assert count(sass, k, "FFMA") == 0, k
        assert count_fused_ffma2(sass, k) == 0 and count(sass, k, "FMUL2") == 0, k
    # axis table build uses the compiler's IEEE division (its internal FFMAs are part of a correctly rounded algorithm)
    assert count(sass, "colorlut_axis_table_kernel", "FMUL") >= 1


def test_rgba64_kernels_only_contain_the_division_fmas(sass):
    for k in ("colorlut_direct_kernelILi1ELb1E", "colorlut_direct_kernelILi2ELb1E"):
        n = count(sass, k, "FFMA")
        assert 0 < n <= 12, (k, n)          # B channel: 2 scalar FMAs x 2 domain variants (R,G: the packed pair below)
        # R and G share one f32x2 division: FMUL2 (its product feeds explicit FMAs only) + 2 FFMA2 per domain variant;
        # no other packed multiply may exist (ptxas would contract it with a packed add into FFMA2)
        assert count_fused_ffma2(sass, k) <= 4 and count(sass, k, "FMUL2") <= 2, k
        assert count(sass, k, "FFMA2") - count_fused_ffma2(sass, k) >= 7, k   # the 7 packed R,G lerp products (mul2_exact)
        assert count(sass, k, "FADD2") >= 8, k   # the packed R,G lerps really are in the binary
        # per-pixel conversions use the 2^23 magic number; the only conversion left is the per-thread `size as f32`
        assert count(sass, k, "F2I") == 0 and count(sass, k, "I2F") == 0 and count(sass, k, "I2FP") <= 2, k
        assert count(sass, k, r"LDG\.E\.ENL2\.256") >= 4 or any("256" in i for n_, b in sass.items() if k in n_ for i in b if i.startswith("LDG")), k


def test_blackwell_native_instructions_present(sass):
    stream = [n for n in sass if "colorlut_memo1d_stream_kernel" in n or "colorlut_memo_stream_kernel" in n]
    assert stream
    for n in stream:
        body = "\n".join(sass[n])
        assert "UBLKCP" in body, n                      # cp.async.bulk (TMA engine) in both directions
        assert "SYNCS.ARRIVE.TRANS64" in body and "SYNCS.PHASECHK" in body, n   # mbarrier expect_tx / try_wait
    assert any("LDG.E.ENL2.256" in i for n, b in sass.items() if "colorlut_direct_kernel" in n for i in b)
    assert any(re.search(r"\bACQBULK|UTMACMDFLUSH|UBLKCP", i) for n in stream for i in sass[n])
