"""world_size-2 (and 3) gloo tests of the multi-GPU plumbing (megagta_b200/shards.py) on the CPU: the four collectives the
library's sharded protocol asks its caller to run (include/mgta_cuda.h mgta_collective), on byte tensors."""
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
