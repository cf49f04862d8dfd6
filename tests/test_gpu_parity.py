"""GPU parity tests: libmgta_cuda.so through its C ABI vs the oracle and the reference-binary goldens."""
import hashlib

import numpy as np
import pytest

from megagta_b200 import cabi
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def run_gpu(rd, k, m, **kw):
    with cabi.Context(k, m, **kw) as ctx:
        ctx.set_reads(rd["seq"], rd["start"], max_len=rd["max_len"])
        res = {}
        if m > 1:
            res["h1"] = ctx.histogram(1)
            res["counting"] = ctx.stage1()
            res["is_solid"] = ctx.get_is_solid()
            res["stats1"] = ctx.stats(1)
        res["h2"] = ctx.histogram(2)
        res["stream"], res["meta"], res["totals"] = ctx.stage2()
        res["stats2"] = ctx.stats(2)
        return res


def check_vs_oracle(rd, k, m, got):
    exp_solid = None
    if m > 1:
        exp_solid, exp_ec, _ = O.stage1(rd, k, m)
        assert np.array_equal(got["h1"], O.s1_hist(rd, k))
        assert np.array_equal(got["counting"], exp_ec)
        n = O.solid_bytes(rd, k)
        assert np.array_equal(got["is_solid"][:n], exp_solid[:n])
    stream, meta, totals = O.stage2(rd, k, m, exp_solid)
    assert np.array_equal(got["h2"], O.s2_hist(rd, k, m, exp_solid if exp_solid is not None else np.zeros(8, np.uint8)))
    assert np.array_equal(got["meta"], meta)
    assert np.array_equal(got["totals"], totals)
    assert got["stream"] == stream


def check_vs_golden(got, g):
    assert len(got["stream"]) == g["stream_bytes"]
    assert O.stream_hash(got["stream"]) == g["stream_hash"]
    assert O.meta_hash(got["meta"]) == g["meta_hash"]
    assert [int(x) for x in got["totals"][:9]] == g["num_w"]
    if g["m"] > 1:
        txt = O.counting_text(got["counting"])
        assert hashlib.sha256(txt.encode()).hexdigest()[:16] == g["counting_sha"]


ORACLE_CASES = [("tiny", 21, 1), ("tiny", 25, 2), ("tiny", 13, 2), ("smoke", 31, 2), ("smoke", 21, 2), ("smoke", 32, 2),
                ("smoke", 41, 3), ("smoke", 61, 2), ("smoke", 99, 2), ("smoke", 63, 1), ("smoke", 127, 2),
                ("adversarial", 31, 2), ("adversarial", 17, 2), ("adversarial", 48, 2), ("adversarial", 21, 1),
                ("xander", 29, 1), ("xander", 44, 2)]


@pytest.mark.parametrize("ds,k,m", ORACLE_CASES)
def test_gpu_matches_oracle(read_lib, ds, k, m):
    _, rd = read_lib(ds)
    check_vs_oracle(rd, k, m, run_gpu(rd, k, m))


@pytest.mark.parametrize("ds,k,m,cap", [("smoke", 31, 2, 256), ("adversarial", 31, 2, 128), ("adversarial", 21, 1, 64),
                                        ("tiny", 25, 2, 64), ("smoke", 61, 2, 64)])
def test_gpu_small_tables_force_the_overflow_pass(read_lib, ds, k, m, cap):
    """A tiny per-tile table limit sends every hash tile through the large-table overflow pass of the counting
    pipeline; a tiny on-chip sort tile makes stage-2 prefix tiles oversize (MSD levels)."""
    _, rd = read_lib(ds)
    got = run_gpu(rd, k, m, sort_items_cap=cap)
    check_vs_oracle(rd, k, m, got)
    if m > 1 and ds != "tiny":
        assert got["stats1"]["msd_levels"] >= 1          # overflow tiles of the counting pass


def test_gpu_small_sort_tiles_force_msd_levels(golden, read_lib, monkeypatch):
    g = golden["cases"]["meta200k_k31_m2"]
    _, rd = read_lib(g["dataset"])
    monkeypatch.setenv("MGTA_S2_PB", "16")               # one prefix tile per lv1 bucket: ~38 items against leaves of 32
    got = run_gpu(rd, g["k"], g["m"], sort_items_cap=64)
    check_vs_golden(got, g)
    assert got["stats2"]["msd_levels"] >= 1


@pytest.mark.parametrize("ds,k,m,budget", [("smoke", 31, 2, 14 << 20), ("adversarial", 27, 3, 24 << 20)])
def test_gpu_small_hbm_budget_forces_batches(read_lib, ds, k, m, budget):
    _, rd = read_lib(ds)
    got = run_gpu(rd, k, m, hbm_budget_bytes=budget)
    check_vs_oracle(rd, k, m, got)
    assert got["stats2"]["n_batches"] > 1 or got["stats1"]["n_batches"] > 1


GOLDEN_CASES = ["smoke_k31_m2", "smoke_k21_m2", "smoke_k22_m2", "smoke_k32_m2", "smoke_k41_m2", "smoke_k61_m2",
                "smoke_k99_m2", "smoke_k31_m1", "smoke_k31_m3", "xander_k29_m1", "xander_k44_m1", "adversarial_k31_m2",
                "adversarial_k21_m1", "adversarial_k27_m3", "adversarial_k30_m2", "adversarial_k48_m2",
                "adversarial_k17_m2", "tiny_k21_m1", "tiny_k25_m2", "meta200k_k31_m2", "meta200k_k61_m2",
                "meta200k_k21_m3", "meta1m_k31_m2"]


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_gpu_matches_reference_golden(case, golden, read_lib):
    g = golden["cases"][case]
    _, rd = read_lib(g["dataset"])
    check_vs_golden(run_gpu(rd, g["k"], g["m"]), g)


def test_gpu_shards_concatenate_to_the_whole_graph(read_lib):
    """world-way bucket sharding on one device: shard streams concatenate to the 1-way stream."""
    _, rd = read_lib("smoke")
    k, m = 31, 2
    whole = run_gpu(rd, k, m)
    for world in (2, 3, 8):
        ctxs = [cabi.Context(k, m, rank=r, world=world) for r in range(world)]
        try:
            solid = None
            ec = np.zeros(65536, dtype=np.int64)
            for c in ctxs:
                c.set_reads(rd["seq"], rd["start"], max_len=rd["max_len"])
                ec += c.stage1()
                s = c.get_is_solid()
                solid = s if solid is None else (solid | s)
            assert np.array_equal(ec, whole["counting"])
            streams, meta = [], np.zeros((65536, 3), dtype=np.int64)
            totals = np.zeros(10, dtype=np.int64)
            for c in ctxs:
                c.set_is_solid(solid)
                st, mt, tt = c.stage2()
                streams.append(st)
                meta += mt
                totals += tt
            assert b"".join(streams) == whole["stream"]
            assert np.array_equal(meta, whole["meta"])
            assert np.array_equal(totals, whole["totals"])
        finally:
            for c in ctxs:
                c.close()


@pytest.mark.parametrize("ds,k,m", [("smoke", 31, 2), ("smoke", 31, 1), ("adversarial", 27, 3), ("meta200k", 61, 2), ("tiny", 25, 2)])
def test_gpu_assist_reads_match_oracle(read_lib, data_dir, ds, k, m):
    """Assist reads (mgta_set_reads n_short_reads < n_reads; reference s1.cpp:104-134, s2.cpp:276,529): they count in
    stage 1, never get is_solid bits, and all their edges are solid in stage 2."""
    import datasets
    _, rd = read_lib(ds)
    rd2, n_short = O.with_assist(rd, datasets.assist_fasta(ds, data_dir))
    exp = O.build_graph(rd2, k, m, False, n_short=n_short)
    with cabi.Context(k, m) as ctx:
        ctx.set_reads(rd2["seq"], rd2["start"], n_short=n_short, max_len=rd2["max_len"])
        if m > 1:
            ec = ctx.stage1()
            assert np.array_equal(ec, exp["counting"])
            n = O.solid_bytes(rd2, k, n_short)
            assert np.array_equal(ctx.get_is_solid()[:n], exp["is_solid"][:n])
        stream, meta, totals = ctx.stage2()
    assert stream == exp["stream"]
    assert np.array_equal(meta, exp["meta"]) and np.array_equal(totals, exp["totals"])


def test_gpu_async_upload_pipelined_with_stage1(golden, read_lib):
    """mgta_set_reads_async from pinned memory: stage 1 extracts chunk c while chunk c + 1 .. are still being copied;
    a second context takes the pageable (synchronous) fallback; both must reproduce the reference golden."""
    import torch
    g = golden["cases"]["meta1m_k31_m2"]
    _, rd = read_lib(g["dataset"])
    seq = torch.from_numpy(np.ascontiguousarray(rd["seq"], dtype=np.uint32).view(np.int32)).pin_memory()
    start = torch.from_numpy(np.ascontiguousarray(rd["start"], dtype=np.uint64).view(np.int64)).pin_memory()
    n = len(rd["start"]) - 1
    for pinned in (True, False):
        with cabi.Context(g["k"], g["m"]) as ctx:
            if pinned:
                ctx.set_reads_async(seq.data_ptr(), seq.numel(), start.data_ptr(), n, n, rd["max_len"])
            else:
                a, b = seq.numpy().copy(), start.numpy().copy()
                ctx.set_reads_async(a.ctypes.data, len(a), b.ctypes.data, n, n, rd["max_len"])
            got = {"counting": ctx.stage1()}
            assert ctx.stats(1)["n_launches"] >= (14 if pinned else 7)      # 8 extraction launches when pipelined
            got["stream"], got["meta"], got["totals"] = ctx.stage2()
            check_vs_golden(got, g)
            if pinned:                                   # a histogram call right after an asynchronous upload waits for it
                ctx.set_reads_async(seq.data_ptr(), seq.numel(), start.data_ptr(), n, n, rd["max_len"])
                assert np.array_equal(ctx.histogram(1), O.s1_hist(rd, g["k"]))


MERCY_CASES = ["tiny_k25_m2_mercy", "smoke_k31_m2_mercy", "smoke_k27_m3_mercy", "xander_k29_m2_mercy",
               "adversarial_k31_m2_mercy", "adversarial_k27_m3_mercy"]


@pytest.mark.parametrize("case", MERCY_CASES)
def test_gpu_mercy_matches_reference_golden_and_oracle(case, golden, read_lib):
    """--need_mercy on the device: candidate multiset (s1.cpp:762-826), "Number mercy" and the extended
    is_solid of s2_read_mercy_prepare (s2.cpp:106-250), and the SdBG built from it, against the goldens the unmodified
    reference produced with --need_mercy and against the oracle."""
    g = golden["cases"][case]
    _, rd = read_lib(g["dataset"])
    k, m = g["k"], g["m"]
    exp = O.build_graph(rd, k, m, True)
    with cabi.Context(k, m, need_mercy=True) as ctx:
        ctx.set_reads(rd["seq"], rd["start"], max_len=rd["max_len"])
        got = {"counting": ctx.stage1()}
        cands = ctx.mercy_candidates()
        assert len(cands) == g["mercy_cand_n"]
        assert hashlib.sha256(np.sort(cands).astype("<u8").tobytes()).hexdigest()[:16] == g["mercy_cand_sha"]
        assert ctx.num_mercy() == g["num_mercy"]
        n = O.solid_bytes(rd, k)
        assert np.array_equal(ctx.get_is_solid()[:n], exp["is_solid"][:n])
        got["stream"], got["meta"], got["totals"] = ctx.stage2()
    check_vs_golden(got, g)
    assert got["stream"] == exp["stream"]


def test_gpu_mercy_with_small_budget_and_bigger_input(golden, read_lib):
    """mercy over several batches of level-1 bins (small HBM budget) on 200k reads: equal to the oracle"""
    _, rd = read_lib("meta200k")
    k, m = 31, 2
    exp = O.build_graph(rd, k, m, True)
    ref_c = O.stage1(rd, k, m, True)[2]
    for budget in (0, 96 << 20):
        with cabi.Context(k, m, need_mercy=True, hbm_budget_bytes=budget) as ctx:
            ctx.set_reads(rd["seq"], rd["start"], max_len=rd["max_len"])
            ctx.stage1()
            assert np.array_equal(np.sort(ctx.mercy_candidates()), np.sort(ref_c))
            assert ctx.num_mercy() == exp["num_mercy"]
            stream, meta, totals = ctx.stage2()
        assert stream == exp["stream"] and np.array_equal(meta, exp["meta"])
