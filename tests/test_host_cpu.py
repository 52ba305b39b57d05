"""CPU tests of the product's host side: the C-ABI library loads, exports every symbol the header
declares, its .cube parser agrees with the reference KATs and with the oracle parser, and it
fails loudly (no CPU fallback) when no CUDA device is present."""
import ctypes as C
import os

import numpy as np
import pytest

import b200vfx
import oracle_binding as orc
from b200vfx import synth


def test_library_exports_every_declared_symbol():
    names = b200vfx.exported_symbols_in_header()
    assert len(names) >= 20
    L = C.CDLL(b200vfx.LIB_PATH)
    for n in names:
        assert hasattr(L, n), "libb200vfx.so does not export %s" % n
    assert b200vfx.lib().b200vfx_abi_version() == 1


def test_library_exports_every_element_layer_symbol():
    import re
    text = open(os.path.join(b200vfx.REPO_ROOT, "include", "b200gst.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = sorted(set(re.findall(r"\b(b200gst_[a-z0-9_]+)\s*\(", text)))
    assert len(names) >= 20
    L = C.CDLL(b200vfx.LIB_PATH)
    for n in names:
        assert hasattr(L, n), "libb200vfx.so does not export %s" % n


def test_no_gpu_fails_loudly():
    if b200vfx.device_count() > 0:
        pytest.skip("GPU present")
    with pytest.raises(b200vfx.B200VfxError) as e:
        b200vfx.Context()
    assert e.value.code == b200vfx.ERR_CUDA and "no CPU fallback" in e.value.msg


# the reference's parser unit tests (video/colorlut/src/parser.rs:382-473) against the PRODUCT parser
def test_product_parser_reference_kats():
    k, s, v, sc, of = b200vfx.cube_parse("\n LUT_3D_SIZE 2\n\n 0.0 0.0 0.0\n 1.0 0.0 0.0\n 0.0 1.0 0.0\n 1.0 1.0 0.0\n"
                                         " 0.0 0.0 1.0\n 1.0 0.0 1.0\n 0.0 1.0 1.0\n 1.0 1.0 1.0\n ")
    assert (k, s) == (3, 2) and v.shape == (8, 3) and v[0].tolist() == [0, 0, 0] and v[7].tolist() == [1, 1, 1]
    k, s, v, sc, of = b200vfx.cube_parse('LUT_1D_SIZE 2\n\nTITLE "test"\nDOMAIN_MIN 0.0 0.0 0.0\nDOMAIN_MAX 1.0 1.0 1.0\n\n'
                                         "0.0 0.0 0.0\n1.0 0.5 0.7\n")
    assert (k, s) == (1, 2) and v[:, 0].tolist() == [0, 1] and v[:, 1].tolist() == [0, 0.5]
    assert v[:, 2].tolist() == [0, np.float32(0.7)]
    for bad in ['LUT_1D_SIZE 2\n0.0 0.0 0.0\n1.0 0.0 0.0\nTITLE "invalid"\n',
                'LUT_1D_SIZE 2\n0.0 0.0 0.0\nTITLE "invalid"\n1.0 0.0 0.0\n',
                "LUT_1D_SIZE 2\nLUT_3D_SIZE 2\n0.0 0.0 0.0\n1.0 1.0 1.0\n"]:
        with pytest.raises(b200vfx.B200VfxError) as e:
            b200vfx.cube_parse(bad)
        assert e.value.code == b200vfx.ERR_PARSE


CASES_OK = [
    synth.cube_text_3d(3, "mix", domain=((-0.5, 0.0, 0.25), (1.0, 2.0, 0.75)), title="t"),
    synth.cube_text_1d(7, 2.2),
    "# c\r\nTITLE \"x y z\"\r\nLUT_1D_SIZE +3\r\nDOMAIN_MIN -1 0 .5\nDOMAIN_MAX 1 2 1.\n1e-1 +.5 5.E-1\n inf -Infinity NaN \n\t0.1 0.2　0.3",
    "LUT_1D_SIZE 2\n0 0 0\n1 1 1\r",
    "LUT_1D_SIZE 2\n\n#x\n   \n1e400 -1e400 1e-60\n0 0 4.9e-324\n",
]
CASES_BAD = [
    "0 0 0\nLUT_1D_SIZE 2\n", "LUT_1D_SIZE 2\n0 0 0\n", "LUT_3D_SIZE 2\n" + "0 0 0\n" * 7, "LUT_1D_SIZE 1\n0 0 0\n",
    "LUT_3D_SIZE 257\n", "LUT_1D_SIZE 65537\n", "LUT_1D_SIZE 2\n0 0\n1 1 1\n", "LUT_1D_SIZE 2\n0 0 0 0\n1 1 1\n",
    "LUT_1D_SIZE 2\n0 0 0x1p0\n1 1 1\n", "LUT_1D_SIZE 2\n0 0 1f\n1 1 1\n", "LUT_1D_SIZE -2\n", "LUT_1D_SIZE 2 3\n",
    "LUT_1D_SIZE\n", "LUT_1D_SIZE 2\nDOMAIN_MIN 1 0 0\nDOMAIN_MAX 1 1 1\n0 0 0\n1 1 1\n",
    "LUT_1D_SIZE 2\nDOMAIN_MIN 0 0\n0 0 0\n1 1 1\n", "LUT_1D_SIZE 2\nDOMAIN_MAX 1 1 1 1\n0 0 0\n1 1 1\n",
    "LUT_1D_SIZE 2\nLUT_3D_INPUT_RANGE 0 1\n0 0 0\n1 1 1\n", "﻿LUT_1D_SIZE 2\n0 0 0\n1 1 1\n", "", "# only\n",
    "LUT_1D_SIZE 2\n0 0 .\n1 1 1\n", "LUT_1D_SIZE 2\n0 0 1e\n1 1 1\n", "LUT_1D_SIZE 2\n0 0 +\n1 1 1\n",
    "LUT_1D_SIZE 99999999999999999999999\n", "title x\nLUT_1D_SIZE 2\n0 0 0\n1 1 1\n",
]


@pytest.mark.parametrize("i", range(len(CASES_OK)))
def test_product_parser_matches_oracle_parser(i):
    text = CASES_OK[i]
    k, s, v, sc, of = b200vfx.cube_parse(text)
    o = orc.cube_parse(text)
    assert (k, s) == (o.kind, o.size)
    assert v.tobytes() == o.values.tobytes() and sc.tobytes() == o.scale.tobytes() and of.tobytes() == o.offset.tobytes()


@pytest.mark.parametrize("i", range(len(CASES_BAD)))
def test_product_parser_rejects_what_oracle_rejects(i):
    with pytest.raises(orc.CubeError):
        orc.cube_parse(CASES_BAD[i])
    with pytest.raises(b200vfx.B200VfxError) as e:
        b200vfx.cube_parse(CASES_BAD[i])
    assert e.value.code == b200vfx.ERR_PARSE


def test_product_parser_invalid_utf8_and_missing_file(tmp_path):
    with pytest.raises(b200vfx.B200VfxError) as e:
        b200vfx.cube_parse(b"LUT_1D_SIZE 2\n0 0 0\n1 1 \xff1\n")
    assert e.value.code == b200vfx.ERR_IO
    L = b200vfx.lib()
    kind, size = C.c_int(), C.c_int()
    vals = C.POINTER(C.c_float)()
    sc = (C.c_float * 3)()
    of = (C.c_float * 3)()
    err = C.create_string_buffer(256)
    rc = L.b200vfx_cube_parse_file(str(tmp_path / "missing.cube").encode(), C.byref(kind), C.byref(size),
                                   C.byref(vals), sc, of, err, 256)
    assert rc == b200vfx.ERR_IO and b"IO error" in err.value
    p = tmp_path / "ok.cube"
    p.write_text(synth.cube_text_3d(2, "identity"))
    rc = L.b200vfx_cube_parse_file(str(p).encode(), C.byref(kind), C.byref(size), C.byref(vals), sc, of, err, 256)
    assert rc == 0 and kind.value == 3 and size.value == 2
    L.b200vfx_cube_free(vals)


def test_product_parser_is_locale_independent():
    """Rust's str::parse::<f32> ignores the process locale; so must the product parser (strtof is LC_NUMERIC-dependent)"""
    import locale
    text = "LUT_1D_SIZE 2\n0.25 0.5 0.75\n1.0 1e0 1.5e-1\n"
    ref = b200vfx.cube_parse(text)
    old = locale.setlocale(locale.LC_NUMERIC)
    tried = []
    try:
        for name in ("de_DE.UTF-8", "de_DE.utf8", "fr_FR.UTF-8", "fr_FR.utf8", "de_DE", "fr_FR", "ru_RU.UTF-8", "nl_NL.UTF-8"):
            try:
                locale.setlocale(locale.LC_NUMERIC, name)
            except locale.Error:
                continue
            tried.append(name)
            got = b200vfx.cube_parse(text)
            assert got[0] == ref[0] and got[1] == ref[1] and np.array_equal(np.asarray(got[2]), np.asarray(ref[2])), name
    finally:
        locale.setlocale(locale.LC_NUMERIC, old)
    # the container may ship no comma-decimal locale: then only the "C" path ran (values checked against the KATs elsewhere)
    assert np.allclose(np.asarray(ref[2]).reshape(-1)[:3], [0.25, 0.5, 0.75])
    if not tried:
        pytest.skip("no comma-decimal locale installed in this image")


def test_blockhash_bits_and_distance_host_side():
    rng = np.random.default_rng(3)
    for _ in range(20):
        sums = rng.integers(0, 2 ** 31, 64, dtype=np.uint32)
        sums[rng.integers(0, 64, 8)] = sums[0]
        assert (b200vfx.blockhash_bits(sums, 3840, 2160) == orc.blockhash_bits(sums, 3840, 2160)).all()
    solid = np.full(64, 255 * 480 * 270, np.uint32)  # solid red frame: every block sum equal
    assert b200vfx.blockhash_bits(solid, 3840, 2160).sum() == 0
    a = b200vfx.blockhash_bits(rng.integers(0, 2 ** 31, 64, dtype=np.uint32), 3840, 2160)
    assert b200vfx.hash_distance(a, a) == 0 and b200vfx.hash_distance(a, 1 - a) == 64


# ---- differential fuzz: product parser (C++) vs oracle parser (C) on generated .cube-like text ------------------------
def _fuzz_strategy():
    from hypothesis import strategies as st
    num = st.one_of(
        st.sampled_from(["0", "1", "0.5", "-0.25", "+1.5", ".5", "5.", "1e-3", "1E+2", "1e400", "-1e-60", "inf", "-inf", "NaN", "nan",
                         "infinity", "+Infinity", "0x10", "1f", "1_0", "1e", "e5", ".", "+", "--1", "1.2.3", "٣", "1e+", "0.1e-0"]),
        st.floats(allow_nan=False, allow_infinity=False, width=32).map(lambda v: repr(float(v))),
        st.integers(-3, 70000).map(str))
    ws = st.sampled_from([" ", "  ", "\t", "　", " \t ", " "])
    data_line = st.lists(num, min_size=1, max_size=5).flatmap(lambda xs: ws.map(lambda w: w.join(xs)))
    kw_line = st.one_of(
        st.tuples(st.sampled_from(["LUT_1D_SIZE", "LUT_3D_SIZE", "lut_1d_size", "LUT_2D_SIZE"]), st.sampled_from(["2", "3", "1", "+2", "-2", "2 2", "", "257", "x", "2.0"])).map(" ".join),
        st.tuples(st.sampled_from(["DOMAIN_MIN", "DOMAIN_MAX"]), st.lists(num, min_size=2, max_size=4).map(" ".join)).map(" ".join),
        st.sampled_from(['TITLE "a b"', "TITLE", "# comment", "", "   ", "LUT_3D_INPUT_RANGE 0 1", "#LUT_1D_SIZE 9"]))
    line = st.one_of(kw_line, data_line, data_line, data_line)
    eol = st.sampled_from(["\n", "\r\n", "\n", "\r", "\n\n"])
    return st.lists(st.tuples(line, eol), min_size=0, max_size=14).map(lambda ls: "".join(a + b for a, b in ls))


def test_product_parser_differential_fuzz():
    hyp = pytest.importorskip("hypothesis")
    from hypothesis import given, settings, HealthCheck

    seen = {"ok": 0, "bad": 0}

    @settings(max_examples=600, deadline=None, derandomize=True, suppress_health_check=list(HealthCheck))
    @given(_fuzz_strategy())
    def run(text):
        try:
            o = orc.cube_parse(text)
        except orc.CubeError:
            o = None
        try:
            p = b200vfx.cube_parse(text)
        except b200vfx.B200VfxError as e:
            assert e.code == b200vfx.ERR_PARSE
            p = None
        assert (o is None) == (p is None), repr(text)
        if o is not None:
            k, s, v, sc, of = p
            assert (k, s) == (o.kind, o.size), repr(text)
            assert v.tobytes() == o.values.tobytes() and sc.tobytes() == o.scale.tobytes() and of.tobytes() == o.offset.tobytes(), repr(text)
            seen["ok"] += 1
        else:
            seen["bad"] += 1

    run()
    assert seen["bad"] > 50   # most random texts are invalid; valid ones are exercised by the structured generator below


def test_product_parser_differential_fuzz_valid_shapes():
    """structured generator: always a well-formed header + the right number of data lines, numbers in odd but legal forms"""
    hyp = pytest.importorskip("hypothesis")
    from hypothesis import given, settings, HealthCheck, strategies as st

    forms = st.sampled_from(["0", "1", "0.5", "-0.25", "+1.5", ".5", "5.", "1e-3", "1E+2", "1e400", "-1e-60", "inf", "-inf", "NaN",
                             "infinity", "0.1e-0", "007", "1e0000000002", "0.30000001192092896", "16777217", "3.4028236e38"])
    n_ok = {"n": 0}

    @settings(max_examples=150, deadline=None, derandomize=True, suppress_health_check=list(HealthCheck))
    @given(st.data())
    def run(data):
        kind = data.draw(st.sampled_from([1, 3]))
        size = data.draw(st.integers(2, 4))
        n = size if kind == 1 else size ** 3
        head = ["LUT_%dD_SIZE %d" % (kind, size)]
        if data.draw(st.booleans()):
            head.insert(data.draw(st.integers(0, 1)), "DOMAIN_MIN %s" % " ".join(data.draw(st.lists(st.sampled_from(["0", "-1", "0.25", "-0.5"]), min_size=3, max_size=3))))
            head.insert(data.draw(st.integers(0, 2)), "DOMAIN_MAX %s" % " ".join(data.draw(st.lists(st.sampled_from(["1", "2", "0.75", "1.5"]), min_size=3, max_size=3))))
        lines = head + [" ".join(data.draw(st.lists(forms, min_size=3, max_size=3))) for _ in range(n)]
        text = data.draw(st.sampled_from(["\n", "\r\n"])).join(lines) + "\n"
        o = orc.cube_parse(text)
        k, s, v, sc, of = b200vfx.cube_parse(text)
        assert (k, s) == (o.kind, o.size) and v.tobytes() == o.values.tobytes(), repr(text)
        assert sc.tobytes() == o.scale.tobytes() and of.tobytes() == o.offset.tobytes(), repr(text)
        n_ok["n"] += 1

    run()
    assert n_ok["n"] >= 100


# ---- the PDL admission rule (host logic) ----------------------------------------------------------------
def _admit(L, key, src, dst, threads, lingers, want=1):
    return L.b200vfx_debug_pdl_admit(key, src[0], src[1], dst[0], dst[1], want, threads, 1 if lingers else 0)


def test_pdl_admission_rule():
    L = b200vfx.lib()
    MB = 1 << 20
    buf = lambda i: (0x10000000 + i * 64 * MB, 0x10000000 + i * 64 * MB + 33 * MB)       # disjoint 33 MB frames
    capped = 148 * 4 * 256            # persistent lookup kernel: 4 CTAs x 256 threads per SM, ends on griddepcontrol.wait
    # 1. capped (lingering) kernels: only the last 2 launches can still be running -> a pool of 3 output buffers is enough
    key = 0x1000
    L.b200vfx_debug_pdl_reset(key)
    got = [_admit(L, key, buf(100 + i), buf(i % 3), capped, True) for i in range(12)]
    assert got == [1] * 12
    # ... a pool of 2 is not: every launch writes what the launch two before it wrote (possibly still running)
    key = 0x1001
    L.b200vfx_debug_pdl_reset(key)
    got = [_admit(L, key, buf(100 + i), buf(i % 2), capped, True) for i in range(8)]
    assert got[:2] == [1, 1] and 0 in got[2:]
    # 2. small frames / non-lingering kernels: nothing bounds how many are co-resident, so every launch since the last
    #    barrier counts as running -> reusing ANY buffer of the chain is refused (and the refusal is the barrier) ...
    small = 300 * 256
    for pool in (4, 5, 9):
        key = 0x1002
        L.b200vfx_debug_pdl_reset(key)
        got = [_admit(L, key, buf(100 + i), buf(i % pool), small, False) for i in range(3 * pool)]
        assert got == [1] * pool + ([0] + [1] * (pool - 1)) * 2, (pool, got)
    #    ... and with all-distinct buffers the bounded record forces a plain launch when it is full (16 launches)
    key = 0x1003
    L.b200vfx_debug_pdl_reset(key)
    got = [_admit(L, key, buf(100 + i), buf(200 + i), small, False) for i in range(40)]
    assert got[:16] == [1] * 16 and got[16] == 0 and got[17:32] == [1] * 15 and got[32] == 0   # the barrier launch is entry 1 of the new record
    #    a lingering stream never fills the record (completed launches are forgotten)
    key = 0x1007
    L.b200vfx_debug_pdl_reset(key)
    assert [_admit(L, key, buf(100 + i), buf(200 + i), capped, True) for i in range(64)] == [1] * 64
    #    one non-lingering launch inside a lingering stream removes the occupancy bound for everything behind it
    key = 0x1008
    L.b200vfx_debug_pdl_reset(key)
    assert _admit(L, key, buf(0), buf(1), capped, True) == 1
    assert _admit(L, key, buf(2), buf(3), small, False) == 1
    assert _admit(L, key, buf(4), buf(5), capped, True) == 1
    assert _admit(L, key, buf(6), buf(7), capped, True) == 1
    assert _admit(L, key, buf(8), buf(1), capped, True) == 0          # 4 launches back, but the chain is not all-lingering
    # 3. chained elements (input = the previous launch's output) and in-place reuse never overlap
    key = 0x1004
    L.b200vfx_debug_pdl_reset(key)
    assert _admit(L, key, buf(0), buf(1), capped, True) == 1
    assert _admit(L, key, buf(1), buf(2), capped, True) == 0          # reads what the previous launch writes
    assert _admit(L, key, buf(5), buf(5), capped, True) == 1          # a refusal is a full barrier: the queue restarts
    assert _admit(L, key, buf(5), buf(5), capped, True) == 0          # same frame in place again
    # 4. a launch that does not ask for PDL is a barrier too; one that is huge (grid >> device) shortens the window to 1
    key = 0x1005
    L.b200vfx_debug_pdl_reset(key)
    assert _admit(L, key, buf(0), buf(1), capped, True) == 1
    assert _admit(L, key, buf(2), buf(3), capped, True, want=0) == 0
    assert _admit(L, key, buf(4), buf(1), capped, True) == 1          # buf(1) was written before the barrier
    huge = 148 * 2048 * 4
    key = 0x1006
    L.b200vfx_debug_pdl_reset(key)
    assert _admit(L, key, buf(0), buf(1), huge, True) == 1
    assert _admit(L, key, buf(2), buf(3), huge, True) == 1
    assert _admit(L, key, buf(4), buf(1), huge, True) == 1            # two launches back: cannot still be resident
    assert _admit(L, key, buf(6), buf(1), huge, True) == 0            # the previous launch itself
    # 5. reductions (block sums, colordetect): the kernel writes only after its griddepcontrol.wait -> rewriting the result
    #    buffer of the launch before is fine, reading a frame that launch is still writing is not; and a later launch
    #    that overwrites the frame a reduction may still be reading is refused
    key = 0x1009
    L.b200vfx_debug_pdl_reset(key)
    red = 148 * 256
    sums = (0x90000000, 0x90000100)
    assert L.b200vfx_debug_pdl_admit(key, *buf(0), *sums, 1, red, 3) == 1
    assert L.b200vfx_debug_pdl_admit(key, *buf(1), *sums, 1, red, 3) == 1      # same result buffer: written after the wait
    assert L.b200vfx_debug_pdl_admit(key, *buf(2), *sums, 1, red, 1) == 0      # an ordinary kernel may not
    assert _admit(L, key, buf(3), buf(4), capped, True) == 1                    # colorlut writes buf(4) ...
    assert L.b200vfx_debug_pdl_admit(key, *buf(4), *sums, 1, red, 3) == 0      # ... a reduction reading it waits
    assert L.b200vfx_debug_pdl_admit(key, *buf(5), *sums, 1, red, 3) == 1
    assert _admit(L, key, buf(6), buf(5), capped, True) == 0                    # overwriting the frame being reduced
    for k in range(0x1000, 0x100A):
        L.b200vfx_debug_pdl_reset(k)
