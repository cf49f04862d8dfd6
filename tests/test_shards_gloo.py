"""world_size-2 (and 3) gloo test of the multi-GPU exchange plumbing (megagta_b200/shards.py) with fake shards on the
CPU: every rank must end up with all rows in rank order and the summed histogram."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from megagta_b200 import shards


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _rows(rank, row_words):
    rng = np.random.default_rng(100 + rank)
    n = [7, 0, 12, 3][rank % 4] if rank != 1 else 5
    return torch.from_numpy(rng.integers(-2**31, 2**31 - 1, size=(n, row_words), dtype=np.int64).astype(np.int32))


def _worker(rank, world, port, row_words, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        local = _rows(rank, row_words)
        hist = torch.full((64,), rank + 1, dtype=torch.int32)
        holder = {}

        def reserve(total, off):
            buf = torch.zeros(max(total, 1) * row_words, dtype=torch.int32)
            buf[off * row_words:(off + len(local)) * row_words] = local.reshape(-1)
            holder["buf"] = buf
            return buf

        counts, offs = shards.exchange(rank, world, dist, torch.device("cpu"), len(local), row_words, reserve, hist)
        exp = torch.cat([_rows(r, row_words) for r in range(world)]).reshape(-1)
        ok = counts == [len(_rows(r, row_words)) for r in range(world)] and offs[-1] * row_words == len(exp) \
            and torch.equal(holder["buf"][:len(exp)], exp) and bool((hist == world * (world + 1) // 2).all())
        out[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_edge_exchange_over_gloo(world):
    mgr = mp.Manager()
    out = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(world, port, 3, out), nprocs=world, join=True)
    assert dict(out) == {r: True for r in range(world)}


def test_plan_offsets():
    assert shards.plan([3, 0, 5]) == [0, 3, 3, 8]


class _FakeCtx:
    """Stands in for cabi.Context in the scan-sharded stage-1 driver: 'items' are int32 words tagged with (source, dest),
    the first scan overflows when `skew` is set (so the agreed rescan path is exercised)."""

    def __init__(self, rank, world, skew):
        self.rank, self.world, self.skew = rank, world, skew
        self.slab_words = 0
        self.scans = 0

    def counts(self):
        return [3 + 2 * self.rank + d + (40 if self.skew and self.rank == 1 and d == 0 else 0) for d in range(self.world)]

    def stage1_slab_items(self):
        return 16

    def stage1_scan(self, lo, hi, slab):
        self.scans += 1
        need = max(max(self.counts()), slab)
        if need > slab:
            return need
        self.slab_words = slab
        self.send = torch.zeros(self.world * slab, dtype=torch.int32)
        for d, c in enumerate(self.counts()):
            self.send[d * slab:d * slab + c] = 1000 * self.rank + 10 * d + 1
        self.recv = torch.full((self.world * slab,), -1, dtype=torch.int32)
        return slab

    def stage1_exchange_buffers(self):
        return self.send, self.recv, self.slab_words * 4, self.counts()

    def stage1_count(self, got):
        self.got = got
        return np.zeros(4, dtype=np.int64)


def _worker_items(rank, world, port, skew, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ctx = _FakeCtx(rank, world, skew)
        orig = shards.DevBuf
        shards.DevBuf = lambda t, nbytes: t            # the fake buffers are tensors already
        as_tensor = torch.as_tensor
        try:
            torch.as_tensor = lambda x, device=None, **kw: x if isinstance(x, torch.Tensor) else as_tensor(x, **kw)
            shards.stage1_scan_sharded(ctx, 1000, rank, world, dist, torch.device("cpu"))
        finally:
            torch.as_tensor = as_tensor
            shards.DevBuf = orig
        slab = ctx.slab_words
        ok = ctx.scans == (2 if skew else 1)
        for s in range(world):
            c = _FakeCtx(s, world, skew).counts()[rank]
            ok = ok and ctx.got[s] == c and bool((ctx.recv[s * slab:s * slab + c] == 1000 * s + 10 * rank + 1).all())
        out[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,skew", [(2, False), (3, False), (2, True)])
def test_stage1_item_exchange_over_gloo(world, skew):
    """scan -> one all-gather of (counts, need) -> agreed rescan on overflow -> equal-split all-to-all -> count"""
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker_items, args=(world, _free_port(), skew, out), nprocs=world, join=True)
    assert dict(out) == {r: True for r in range(world)}


def test_read_ranges_partition_the_reads():
    for n, w in [(10, 3), (1000, 7), (5, 8)]:
        r = [shards.read_range(n, i, w) for i in range(w)]
        assert r[0][0] == 0 and r[-1][1] == n and all(r[i][1] == r[i + 1][0] for i in range(w - 1))


# ---- the library-driven protocol: the four collectives of include/mgta_cuda.h on byte tensors ---------------------------
def _coll_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from megagta_b200 import cabi
        n = 24                                                   # bytes per peer
        ok = True
        # ALL_TO_ALL: slab d of send goes to rank d, slab s of recv comes from rank s
        send = torch.tensor([(rank * 16 + d) for d in range(world) for _ in range(n)], dtype=torch.uint8)
        recv = torch.zeros(world * n, dtype=torch.uint8)
        shards.run_collective_tensors(cabi.COLL_ALL_TO_ALL, send, recv, rank, world, dist)
        ok &= recv.tolist() == [(s * 16 + rank) for s in range(world) for _ in range(n)]
        # ALL_GATHER in place: my part already sits at my slot
        buf = torch.zeros(world * n, dtype=torch.uint8)
        buf[rank * n:(rank + 1) * n] = rank + 1
        shards.run_collective_tensors(cabi.COLL_ALL_GATHER, buf[rank * n:(rank + 1) * n], buf, rank, world, dist)
        ok &= buf.tolist() == [s + 1 for s in range(world) for _ in range(n)]
        # ALL_REDUCE SUM u32: wraps like unsigned arithmetic (summed as int32 words)
        a = np.array([0xF0000000 + rank, 7, 0xFFFFFFFF], dtype=np.uint32)
        t = torch.from_numpy(a.view(np.uint8).copy())
        shards.run_collective_tensors(cabi.COLL_ALL_REDUCE_SUM_U32, t, t, rank, world, dist)
        exp = np.array([(sum(0xF0000000 + r for r in range(world))) & 0xFFFFFFFF, 7 * world, (0xFFFFFFFF * world) & 0xFFFFFFFF], dtype=np.uint32)
        ok &= np.array_equal(t.numpy().view(np.uint32), exp)
        b = np.array([2**62 + rank, 5], dtype=np.uint64)
        t = torch.from_numpy(b.view(np.uint8).copy())
        shards.run_collective_tensors(cabi.COLL_ALL_REDUCE_SUM_U64, t, t, rank, world, dist)
        exp = np.array([(sum(2**62 + r for r in range(world))) % 2**64, 5 * world], dtype=np.uint64)
        ok &= np.array_equal(t.numpy().view(np.uint64), exp)
        out[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_library_collectives_over_gloo(world):
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_coll_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert dict(out) == {r: True for r in range(world)}
