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
