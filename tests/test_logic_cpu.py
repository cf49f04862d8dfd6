"""CPU tests of the product's __host__ __device__ item/emission logic (kmer_ops.cuh, cx1_items.cuh,
cx1_emit.cuh) driven by tests/cpu/logic_host.cpp, against the oracle.  Sorting is std::sort here; the
device sort is covered by the gpu tests."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle as O
import oracle_memo as OM

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpu", "logic_host.cpp")
OUT = os.path.join(ROOT, "tests", "cpu", "_build", "liblogic_host.so")


@pytest.fixture(scope="session")
def logic():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    deps = [SRC] + [os.path.join(ROOT, "megagta_b200", "csrc", f) for f in ("kmer_ops.cuh", "cx1_items.cuh", "cx1_emit.cuh")]
    if not os.path.exists(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in deps):
        subprocess.run(["g++", "-O2", "-std=c++17", "-x", "c++", "-shared", "-fPIC", "-Wno-unknown-pragmas", "-o", OUT, SRC],
                       check=True)
    return ctypes.CDLL(OUT)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def run_logic(lib, rd, k, m, mercy, n_short=None):
    n, ns = ctypes.c_int64(rd["n_reads"]), ctypes.c_int64(rd["n_reads"] if n_short is None else n_short)
    is_solid = np.zeros(O.solid_bytes(rd, k, n_short) + 8, dtype=np.uint8)
    ec = np.zeros(65536, dtype=np.int64)
    h1 = np.zeros(65536, dtype=np.int64)
    cands = np.empty(0, dtype=np.uint64)
    if m > 1:
        cp, cn = ctypes.c_void_p(), ctypes.c_int64()
        assert lib.logic_stage1(_p(rd["seq"]), _p(rd["start"]), n, ns, rd["max_len"], k, m, _p(is_solid), _p(ec),
                                int(mercy), ctypes.byref(cp), ctypes.byref(cn), _p(h1)) == 0
        cands = np.empty(cn.value, dtype=np.uint64)
        if cn.value:
            ctypes.memmove(_p(cands), cp, cn.value * 8)
        lib.logic_free(cp)
        if mercy:
            O.mercy(rd, k, is_solid, cands, n_short)
    sp, sn = ctypes.c_void_p(), ctypes.c_int64()
    meta = np.zeros((65536, 3), dtype=np.int64)
    totals = np.zeros(10, dtype=np.int64)
    h2 = np.zeros(65536, dtype=np.int64)
    assert lib.logic_stage2(_p(rd["seq"]), _p(rd["start"]), n, ns, rd["max_len"], k, m, _p(is_solid), ctypes.byref(sp),
                            ctypes.byref(sn), _p(meta), _p(totals), _p(h2)) == 0
    stream = ctypes.string_at(sp, sn.value)
    lib.logic_free(sp)
    return dict(stream=stream, meta=meta, totals=totals, counting=ec, is_solid=is_solid, cands=cands, h1=h1, h2=h2)


CASES = [("tiny", 21, 1, False), ("tiny", 25, 2, True), ("tiny", 13, 2, False),
         ("smoke", 31, 2, False), ("smoke", 31, 2, True), ("smoke", 21, 2, False), ("smoke", 32, 2, False),
         ("smoke", 41, 3, True), ("smoke", 61, 2, False), ("smoke", 99, 2, False), ("smoke", 63, 1, False),
         ("adversarial", 31, 2, True), ("adversarial", 17, 2, False), ("adversarial", 48, 2, False),
         ("adversarial", 27, 3, True), ("adversarial", 64, 2, True), ("adversarial", 21, 1, False)]


@pytest.mark.parametrize("ds,k,m,mercy", CASES)
def test_device_logic_on_cpu_matches_oracle(logic, read_lib, ds, k, m, mercy):
    _, rd = read_lib(ds)
    got = run_logic(logic, rd, k, m, mercy)
    exp_solid = exp_ec = None
    exp_cands = np.empty(0, dtype=np.uint64)
    if m > 1:
        exp_solid, exp_ec, exp_cands = OM.stage1(rd, k, m, mercy)
        assert np.array_equal(got["h1"], O.s1_hist(rd, k))
        assert np.array_equal(got["counting"], exp_ec)
        assert np.array_equal(got["cands"], exp_cands)
        if mercy:
            O.mercy(rd, k, exp_solid, exp_cands)
        assert np.array_equal(got["is_solid"][:len(exp_solid)], exp_solid)
    stream, meta, totals = OM.stage2(rd, k, m, exp_solid)
    assert np.array_equal(got["h2"], O.s2_hist(rd, k, m, exp_solid if exp_solid is not None else np.zeros(8, np.uint8)))
    assert got["stream"] == stream
    assert np.array_equal(got["meta"], meta)
    assert np.array_equal(got["totals"], totals)


# ---- edge-centric path (v2): the canonical-(k+1)-mer multiset gives stage 1, and {(edge, solid occurrences)} gives stage 2
def run_edges(lib, rd, k, m, is_solid=None, fused=False, n_short=None):
    n, ns = ctypes.c_int64(rd["n_reads"]), ctypes.c_int64(rd["n_reads"] if n_short is None else n_short)
    sp, sn = ctypes.c_void_p(), ctypes.c_int64()
    meta = np.zeros((65536, 3), dtype=np.int64)
    totals = np.zeros(10, dtype=np.int64)
    sol = is_solid if is_solid is not None else np.zeros(8, dtype=np.uint8)
    assert lib.logic_stage2_edges(_p(rd["seq"]), _p(rd["start"]), n, ns, rd["max_len"], k, m, _p(sol), int(fused),
                                  ctypes.byref(sp), ctypes.byref(sn), _p(meta), _p(totals)) == 0
    stream = ctypes.string_at(sp, sn.value)
    lib.logic_free(sp)
    return stream, meta, totals


EDGE_CASES = [("tiny", 21, 1, False), ("tiny", 25, 2, True), ("tiny", 13, 2, False), ("tiny", 9, 2, False),
              ("smoke", 31, 2, False), ("smoke", 31, 2, True), ("smoke", 21, 2, False), ("smoke", 32, 2, False),
              ("smoke", 41, 3, True), ("smoke", 61, 2, False), ("smoke", 99, 2, False), ("smoke", 63, 1, False),
              ("smoke", 15, 2, False), ("smoke", 47, 2, False),
              ("adversarial", 31, 2, True), ("adversarial", 17, 2, False), ("adversarial", 48, 2, False),
              ("adversarial", 27, 3, True), ("adversarial", 64, 2, True), ("adversarial", 21, 1, False),
              ("xander", 29, 1, False), ("xander", 44, 2, True),
              # the ends of the k range (MGTA_MAX_K = 127, bucket prefix needs k >= 9) and word-boundary sizes of 8-word k-mers
              ("xander", 127, 1, False), ("xander", 126, 2, False), ("xander", 112, 2, True), ("xander", 97, 3, True),
              ("tiny", 10, 2, False), ("tiny", 11, 1, False)]


@pytest.mark.parametrize("ds,k,m,mercy", EDGE_CASES)
def test_edge_centric_logic_matches_oracle(logic, read_lib, ds, k, m, mercy):
    _, rd = read_lib(ds)
    n, ns = ctypes.c_int64(rd["n_reads"]), ctypes.c_int64(rd["n_reads"])
    exp_solid = None
    if m > 1:
        exp_solid, exp_ec, cands = OM.stage1(rd, k, m, mercy)
        got_solid = np.zeros(len(exp_solid) + 8, dtype=np.uint8)
        got_ec = np.zeros(65536, dtype=np.int64)
        assert lib_stage1_edges(logic, rd, k, m, got_solid, got_ec) == 0
        assert np.array_equal(got_ec, exp_ec)
        assert np.array_equal(got_solid[:len(exp_solid)], exp_solid)
        if mercy:
            O.mercy(rd, k, exp_solid, cands)
    exp = OM.stage2(rd, k, m, exp_solid)
    # general mode: re-count under the (possibly mercy-extended) is_solid filter
    stream, meta, totals = run_edges(logic, rd, k, m, exp_solid, fused=False)
    assert stream == exp[0] and np.array_equal(meta, exp[1]) and np.array_equal(totals, exp[2])
    if not mercy:
        # fused mode: multiplicities straight from the stage-1 counts
        stream, meta, totals = run_edges(logic, rd, k, m, None, fused=True)
        assert stream == exp[0] and np.array_equal(meta, exp[1]) and np.array_equal(totals, exp[2])


def lib_stage1_edges(lib, rd, k, m, solid, ec, n_short=None):
    n, ns = ctypes.c_int64(rd["n_reads"]), ctypes.c_int64(rd["n_reads"] if n_short is None else n_short)
    return lib.logic_stage1_edges(_p(rd["seq"]), _p(rd["start"]), n, ns, rd["max_len"], k, m, _p(solid), _p(ec))


MERCY_SCAN_CASES = [("tiny", 25, 2), ("smoke", 31, 2), ("smoke", 27, 3), ("smoke", 41, 3), ("adversarial", 31, 2),
                    ("adversarial", 27, 3), ("adversarial", 64, 2), ("xander", 29, 2), ("xander", 44, 2)]


@pytest.mark.parametrize("ds,k,m", MERCY_SCAN_CASES)
def test_device_mercy_scan_on_cpu_matches_oracle(logic, read_lib, ds, k, m):
    """cx1_emit.cuh mercy_scan_read (the code k_mercy_reads runs) over position flag vectors vs the oracle's restatement of
    s2_read_mercy_prepare over sorted candidates: same is_solid, same "Number mercy"."""
    _, rd = read_lib(ds)
    solid, _, cands = OM.stage1(rd, k, m, True)
    exp = solid.copy()
    exp_n = O.mercy(rd, k, exp, cands)
    got = solid.copy()
    shuffled = np.random.default_rng(5).permutation(cands)          # the device list is unordered and may hold duplicates
    shuffled = np.concatenate([shuffled, shuffled[: len(shuffled) // 3]]).astype(np.uint64)
    logic.logic_mercy.restype = ctypes.c_int64
    n = logic.logic_mercy(_p(np.ascontiguousarray(rd["start"], dtype=np.uint64)), ctypes.c_int64(rd["n_reads"]), rd["max_len"], k,
                          _p(shuffled), ctypes.c_int64(len(shuffled)), _p(got))
    assert n == exp_n
    assert np.array_equal(got, exp)


@pytest.mark.parametrize("ds,k,m,mercy", [c for c in EDGE_CASES if c[0] != "smoke" or c[1:] in ((31, 2, False), (31, 2, True), (63, 1, False))])
def test_node_filter_drops_only_items_the_group_logic_drops(logic, read_lib, ds, k, m, mercy):
    """Design check for the next step: with a k-mer node table (which k-mers a solid edge enters /
    leaves) the $-items that output_() would drop anyway are never generated; the records must not change."""
    _, rd = read_lib(ds)
    exp_solid = None
    if m > 1:
        exp_solid, _, cands = OM.stage1(rd, k, m, mercy)
        if mercy:
            O.mercy(rd, k, exp_solid, cands)
    exp = OM.stage2(rd, k, m, exp_solid)
    logic.logic_last_items.restype = ctypes.c_int64
    run_edges(logic, rd, k, m, exp_solid, fused=0)
    n_all = logic.logic_last_items()
    stream, meta, totals = run_edges(logic, rd, k, m, exp_solid, fused=2)
    n_filtered = logic.logic_last_items()
    assert stream == exp[0] and np.array_equal(meta, exp[1]) and np.array_equal(totals, exp[2])
    assert n_filtered <= n_all
    if ds == "smoke" and k == 31 and m == 2 and not mercy:
        assert n_filtered < 0.5 * n_all                  # two thirds of the items are $-items, almost all of them covered


@pytest.mark.parametrize("ds,k,m,mercy", EDGE_CASES + [("adversarial", 16, 2, False), ("adversarial", 32, 1, False), ("smoke", 30, 2, False)])
def test_node_pass_reproduces_the_records(logic, read_lib, ds, k, m, mercy):
    """The product's stage-2 formulation: real items from the distinct solid edges, $-items only
    for the tip k-mers found by accumulating out / in weights per canonical k-mer (node_ops_of_edge, s2_tip_items --
    the code k_node_part / k_node_count run).  Records, per-bucket table and totals must equal the oracle's; even k
    exercises palindromic k-mers, k % 16 == 0 the word-boundary cases."""
    _, rd = read_lib(ds)
    exp_solid = None
    if m > 1:
        exp_solid, _, cands = OM.stage1(rd, k, m, mercy)
        if mercy:
            O.mercy(rd, k, exp_solid, cands)
    exp = OM.stage2(rd, k, m, exp_solid)
    logic.logic_last_items.restype = ctypes.c_int64
    stream, meta, totals = run_edges(logic, rd, k, m, exp_solid, fused=4)
    assert stream == exp[0] and np.array_equal(meta, exp[1]) and np.array_equal(totals, exp[2])
    assert logic.logic_last_items() <= 2 * int(exp[1][:, 0].sum()) + 2
    if not mercy:
        stream, meta, totals = run_edges(logic, rd, k, m, None, fused=5)     # multiplicities straight from the stage-1 counts
        assert stream == exp[0] and np.array_equal(meta, exp[1]) and np.array_equal(totals, exp[2])


@pytest.mark.parametrize("seed", range(52))
def test_product_logic_on_random_inputs(logic, tmp_path, seed):
    """the seeded random read sets of tests/test_oracle_fuzz.py (ragged lengths, repeats, palindromes, k = 9 ... 127 around
    every word boundary, min-count 1..3, mercy, assist reads) through the product's item / node-pass / emission logic, against
    the oracle (which that file pins on the reference binary run live)"""
    import test_oracle_fuzz as F
    prefix, k, m, mercy, fa = F.make_case(seed, str(tmp_path))
    rd, ns = O.load_read_lib(prefix), None
    if fa:
        rd, ns = O.with_assist(rd, fa)
    exp_solid = None
    if m > 1:
        exp_solid, exp_ec, cands = O.stage1(rd, k, m, mercy, ns)
        got_solid = np.zeros(len(exp_solid) + 8, dtype=np.uint8)
        got_ec = np.zeros(65536, dtype=np.int64)
        assert lib_stage1_edges(logic, rd, k, m, got_solid, got_ec, ns) == 0             # canonical (k+1)-mer counting
        assert np.array_equal(got_ec, exp_ec) and np.array_equal(got_solid[:len(exp_solid)], exp_solid)
        if mercy:
            got = run_logic(logic, rd, k, m, True, ns)                                     # the reference's own stage-1 items + mercy
            assert np.array_equal(got["cands"], cands)
            O.mercy(rd, k, exp_solid, cands, ns)
            assert np.array_equal(got["is_solid"][:len(exp_solid)], exp_solid)
    exp = O.stage2(rd, k, m, exp_solid, ns)
    stream, meta, totals = run_edges(logic, rd, k, m, exp_solid, fused=4, n_short=ns)      # node pass (the product's stage 2)
    assert stream == exp[0] and np.array_equal(meta, exp[1]) and np.array_equal(totals, exp[2])
    if not mercy:
        stream, meta, totals = run_edges(logic, rd, k, m, None, fused=5, n_short=ns)       # multiplicities from the stage-1 counts
        assert stream == exp[0] and np.array_equal(meta, exp[1]) and np.array_equal(totals, exp[2])
