"""The sharded build protocol (mgta_sharded_begin / _step / _result) with `world` contexts on ONE device: the test plays the
role of the caller and runs every collective the library asks for with plain device copies, all shards in lockstep.
What real NCCL does between GPUs (tests/gpu_multi.py under torchrun) is here a loop over contexts; the library code is
the same.  Expected: shard streams concatenate to the one-shard stream, tables / totals add up, edge_counting is whole
on every shard."""
import numpy as np
import pytest

from megagta_b200 import cabi
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _view(ptr, nbytes, dtype="|u1", itemsize=1):
    import torch

    class _Buf:
        __cuda_array_interface__ = {"shape": (nbytes // itemsize,), "typestr": dtype, "data": (ptr, False), "version": 3}
    return torch.as_tensor(_Buf(), device="cuda:0")


def perform(cs):
    """run one collective among the contexts (cs[r] = what shard r was told to do)"""
    import torch
    world, op, n = len(cs), cs[0].op, cs[0].bytes
    assert len({c.op for c in cs}) == 1 and len({c.bytes for c in cs}) == 1
    if op == cabi.COLL_ALL_TO_ALL:
        send = [_view(c.send, world * n) for c in cs]
        recv = [_view(c.recv, world * n) for c in cs]
        for d in range(world):
            for s in range(world):
                recv[d][s * n:(s + 1) * n] = send[s][d * n:(d + 1) * n]
    elif op == cabi.COLL_ALL_GATHER:
        recv = [_view(c.recv, world * n) for c in cs]
        for r, c in enumerate(cs):
            assert c.send == c.recv + r * n                  # in place: my part sits at my slot
        parts = [recv[s][s * n:(s + 1) * n].clone() for s in range(world)]
        for d in range(world):
            for s in range(world):
                recv[d][s * n:(s + 1) * n] = parts[s]
    elif op in (cabi.COLL_ALL_REDUCE_SUM_U32, cabi.COLL_ALL_REDUCE_SUM_U64):
        dt, isz, tt = ("<i4", 4, torch.int32) if op == cabi.COLL_ALL_REDUCE_SUM_U32 else ("<i8", 8, torch.int64)
        bufs = [_view(c.recv, n, dt, isz) for c in cs]
        total = torch.stack(bufs).sum(dim=0, dtype=tt)
        for b in bufs:
            b.copy_(total)
    else:
        raise AssertionError("unknown collective %d" % op)


def lockstep(ctxs, stage, collect=True):
    import torch
    for c in ctxs:
        c.sharded_begin(stage, collect)
    n_coll = 0
    while True:
        cs = [c.sharded_step() for c in ctxs]
        torch.cuda.synchronize()
        if all(x is None for x in cs):
            break
        assert all(x is not None for x in cs), "the shards disagree on whether the stage has finished"
        perform(cs)
        torch.cuda.synchronize()
        n_coll += 1
    return [c.sharded_result() for c in ctxs], n_coll


def one_shard(rd, k, m, n_short=None):
    with cabi.Context(k, m) as ctx:
        ctx.set_reads(rd["seq"], rd["start"], n_short=n_short, max_len=rd["max_len"])
        ec = ctx.stage1() if m > 1 else None
        stream, meta, totals = ctx.stage2()
    return ec, stream, meta, totals


def check_world(rd, k, m, world, whole, n_short=None, **kw):
    ec0, stream0, meta0, totals0 = whole
    ctxs = [cabi.Context(k, m, rank=r, world=world, **kw) for r in range(world)]
    try:
        for c in ctxs:
            c.set_reads(rd["seq"], rd["start"], n_short=n_short, max_len=rd["max_len"])
        if m > 1:
            ecs, n1 = lockstep(ctxs, 1)
            for ec in ecs:
                assert np.array_equal(ec, ec0)                # all-reduced: whole on every shard
            assert n1 >= 1
        res, n2 = lockstep(ctxs, 2)
        assert b"".join(r[0] for r in res) == stream0
        assert np.array_equal(sum(r[1] for r in res), meta0)
        assert np.array_equal(sum(r[2] for r in res), totals0)
        again, _ = lockstep(ctxs, 2)                          # stage 2 once more on the exchanged state: same records
        assert b"".join(r[0] for r in again) == stream0
        return ctxs[0].stats(1), ctxs[0].stats(2)
    finally:
        for c in ctxs:
            c.close()


@pytest.mark.parametrize("ds,k,m", [("smoke", 31, 2), ("adversarial", 27, 3), ("meta200k", 61, 2), ("smoke", 31, 1), ("xander", 44, 2),
                                    ("tiny", 25, 2)])
def test_sharded_protocol_on_one_device(read_lib, ds, k, m):
    _, rd = read_lib(ds)
    whole = one_shard(rd, k, m)
    for world in (2, 3):
        check_world(rd, k, m, world, whole)


def test_sharded_protocol_with_small_tables_and_budget(read_lib):
    """a tiny table limit (overflow passes) and a small HBM budget (several batches)"""
    _, rd = read_lib("meta200k")
    whole = one_shard(rd, 31, 2)
    check_world(rd, 31, 2, 2, whole, sort_items_cap=96)
    check_world(rd, 31, 2, 4, whole, hbm_budget_bytes=200 << 20)


def test_sharded_protocol_with_overflowing_send_slabs(read_lib, monkeypatch):
    """send slabs of 64 items (MGTA_TEST_SLAB_ITEMS, the test hook of the stage-1 and node-pass scans) overflow on every
    shard: the all-gathered table reports the size that fits and every shard rescans once"""
    _, rd = read_lib("smoke")
    whole = one_shard(rd, 31, 2)
    monkeypatch.setenv("MGTA_TEST_SLAB_ITEMS", "64")
    s1, s2 = check_world(rd, 31, 2, 3, whole)
    assert s1["n_giants"] >= 1                                  # stage 1 counted its rescan


def test_sharded_stage1_in_rounds(read_lib, monkeypatch):
    """inputs whose stage-1 items do not fit the HBM at once are exchanged in rounds (one slice of every shard's hash range
    per round, the reads scanned again each time); MGTA_TEST_ROUNDS forces three rounds on a small input"""
    _, rd = read_lib("meta200k")
    whole = one_shard(rd, 31, 2)
    monkeypatch.setenv("MGTA_TEST_ROUNDS", "3")
    s1, _ = check_world(rd, 31, 2, 2, whole)
    assert s1["n_batches"] >= 3
    monkeypatch.setenv("MGTA_TEST_ROUNDS", "5")
    check_world(rd, 61, 2, 3, one_shard(rd, 61, 2))


def test_sharded_protocol_with_assist_reads(read_lib, data_dir):
    import datasets
    _, rd = read_lib("smoke")
    rd2, n_short = O.with_assist(rd, datasets.assist_fasta("smoke", data_dir))
    whole = one_shard(rd2, 31, 2, n_short=n_short)
    check_world(rd2, 31, 2, 2, whole, n_short=n_short)


def test_one_shard_through_the_sharded_entry_points(read_lib):
    """world == 1: the same loop, no collective is ever requested"""
    _, rd = read_lib("smoke")
    whole = one_shard(rd, 31, 2)
    with cabi.Context(31, 2) as ctx:
        ctx.set_reads(rd["seq"], rd["start"], max_len=rd["max_len"])
        ec = ctx.sharded(1, lambda c: pytest.fail("no collective expected"))
        stream, meta, totals = ctx.sharded(2, lambda c: pytest.fail("no collective expected"))
    assert np.array_equal(ec, whole[0]) and stream == whole[1] and np.array_equal(meta, whole[2])


@pytest.mark.parametrize("ds,k,m", [("smoke", 31, 2), ("adversarial", 27, 3), ("xander", 29, 2), ("meta200k", 31, 2)])
def test_sharded_mercy(read_lib, golden, ds, k, m):
    """--need_mercy on several shards: is_solid summed over the shards, candidates of every shard's slice of the (k-1)-mer
    hash space all-gathered, per-read scan on every shard.  Same candidates, "Number mercy", is_solid and graph as one shard
    (which the mercy goldens of the unmodified reference pin, tests/test_gpu_parity.py)."""
    _, rd = read_lib(ds)
    with cabi.Context(k, m, need_mercy=True) as ctx:
        ctx.set_reads(rd["seq"], rd["start"], max_len=rd["max_len"])
        ec0 = ctx.stage1()
        cand0, nm0, solid0 = np.sort(ctx.mercy_candidates()), ctx.num_mercy(), ctx.get_is_solid()
        stream0, meta0, totals0 = ctx.stage2()
    g = golden["cases"].get("%s_k%d_m%d_mercy" % (ds, k, m))
    if g:
        assert nm0 == g["num_mercy"] and len(cand0) == g["mercy_cand_n"] and O.stream_hash(stream0) == g["stream_hash"]
    for world in (2, 3):
        ctxs = [cabi.Context(k, m, need_mercy=True, rank=r, world=world) for r in range(world)]
        try:
            for c in ctxs:
                c.set_reads(rd["seq"], rd["start"], max_len=rd["max_len"])
            ecs, n1 = lockstep(ctxs, 1)
            assert n1 >= 4                                     # item exchange, edge_counting, is_solid, candidate counts (+ candidates)
            for c, ec in zip(ctxs, ecs):
                assert np.array_equal(ec, ec0)
                assert c.num_mercy() == nm0
                assert np.array_equal(np.sort(c.mercy_candidates()), cand0)
                assert np.array_equal(c.get_is_solid(), solid0)
            res, _ = lockstep(ctxs, 2)
            assert b"".join(r[0] for r in res) == stream0
            assert np.array_equal(sum(r[1] for r in res), meta0)
            assert np.array_equal(sum(r[2] for r in res), totals0)
        finally:
            for c in ctxs:
                c.close()
