"""Exchange step between the stages when the graph is sharded over several GPUs.

Stage 1 of shard r leaves the solid edges of ITS hash range on its device; stage 2 of every shard needs all of them
(it emits a bucket range, and an edge's stage-2 items fall into unrelated buckets).  The plumbing is
`torch.distributed` (NCCL over NVLink on the GPUs, gloo in the CPU tests):

    counts    all-reduce of a one-hot int64 vector          -> rows per shard, offsets
    rows      one broadcast per shard into the common buffer (ncclBroadcast of the owner's slice)
    histogram all-reduce SUM of the stage-2 key-prefix histogram (int32 words)

The functions only see tensors, so the CPU tests drive them with fake shards."""
import torch


class DevBuf:
    """int32 view of raw device memory for torch.as_tensor (CUDA array interface)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes // 4,), "typestr": "<i4", "data": (ptr, False), "version": 3}


def plan(counts):
    """rows per shard -> offsets (len world + 1) of the shards' slices in the common row buffer"""
    offs = [0]
    for c in counts:
        offs.append(offs[-1] + int(c))
    return offs


def gather_counts(n_local, rank, world, dist, device):
    counts = torch.zeros(world, dtype=torch.int64, device=device)
    counts[rank] = n_local
    dist.all_reduce(counts)
    return [int(x) for x in counts.tolist()]


def exchange(rank, world, dist, device, n_local, row_words, reserve, hist):
    """reserve(total_rows, my_offset_rows) -> int32 tensor of total_rows * row_words words that already holds the local
    rows at my_offset; hist: int32 tensor (summed in place).  Returns (counts, offsets)."""
    counts = gather_counts(n_local, rank, world, dist, device)
    offs = plan(counts)
    buf = reserve(offs[-1], offs[rank])
    for j in range(world):
        if counts[j]:
            dist.broadcast(buf[offs[j] * row_words:offs[j + 1] * row_words], j)
    dist.all_reduce(hist, op=dist.ReduceOp.SUM)
    return counts, offs


def exchange_ctx(ctx, rank, world, dist, device):
    """The same over a cabi.Context (device pointers from the C ABI)."""
    _, n, w = ctx.edges_local()

    def reserve(total, off):
        p = ctx.edges_reserve(total, off)
        return torch.as_tensor(DevBuf(p, max(total, 1) * w * 4), device=device)

    hp, hb = ctx.edge_hist_device_buffer()
    return exchange(rank, world, dist, device, n, w, reserve, torch.as_tensor(DevBuf(hp, hb), device=device))
