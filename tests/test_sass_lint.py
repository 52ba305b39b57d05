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


def test_round2_kernels_use_the_instructions_their_design_rests_on(sass):
    """cheap guards against a silent de-optimisation by a compiler or source change (the kernels' results would stay right)"""
    def ops(name_part):
        return [i for n, b in sass.items() if name_part in n for i in b]
    # programmatic dependent launch: the reductions trigger their dependents at entry (PREEXIT) ...
    for k in ("blockhash_rows_kernel", "blockhash_sums_kernel", "colordetect_hist_kernel", "colorlut_memo_apply_kernel"):
        assert any(re.match(r"(@!?U?P\d+\s+)?PREEXIT\b", i) for i in ops(k)), k
    # ... the row-streaming block sums merge a warp's columns with match.any + redux before the shared atomic
    rows = ops("blockhash_rows_kernel")
    assert any("MATCH.ANY" in i for i in rows) and any("REDUX.SUM" in i for i in rows) and any("ATOMS.ADD" in i for i in rows)
    # luma and the RGB -> YUV matrix are dp4a on packed bytes (unsigned x unsigned for Y, unsigned x signed for chroma)
    assert any("IDP.4A.U8.U8" in i for i in ops("luma_vresize_kernel"))
    planar = ops("colorlut_i420_x8_kernel")
    assert any("IDP.4A.U8.U8" in i for i in planar) and any("IDP.4A.U8.S8" in i for i in planar)
    # the horizontal resize chain reads its products 16 bytes at a time and adds them with dependent FADDs (never FFMA)
    hres = ops("luma_hresize_kernel")
    assert any("LDS.128" in i for i in hres) and not any(re.match(r"(@!?U?P\d+\s+)?FFMA\b", i) for i in hres)
    assert not any(re.match(r"(@!?U?P\d+\s+)?FFMA\b", i) for i in ops("luma_vresize_kernel"))
    # 32-byte stores of the fused tile gather, the multicast store path compiles to system-scope 16-byte stores
    tg = ops("colorlut_tile_gather_kernel")
    assert any("STG.E.ENL2.256" in i for i in tg) and any("STG.E.128.STRONG.SYS" in i for i in tg)
