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


def read_range(n_reads, rank, world):
    """reads [begin, end) that shard `rank` scans in the scan-sharded stage 1: equal contiguous slices"""
    return n_reads * rank // world, n_reads * (rank + 1) // world


def agree_max(value, dist, device):
    t = torch.as_tensor([int(value)], dtype=torch.int64).to(device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return int(t.item())


def share_counts(rank, world, dist, device, send_counts, need):
    """One all-gather tells every shard what every shard sends to whom, and the largest slab any shard needs.
    -> (items this shard receives from each shard, max need)"""
    mine = torch.as_tensor([int(x) for x in send_counts] + [int(need)], dtype=torch.int64).to(device)
    table = torch.empty(world * (world + 1), dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(table, mine)
    t = table.view(world, world + 1).cpu()
    return [int(t[s, rank]) for s in range(world)], int(t[:, world].max())


def exchange_items(rank, world, dist, device, send, recv):
    """All-to-all of the stage-1 items between the scan and the count step.  send / recv: int32 tensors of `world`
    equal slabs (slab d of `send` goes to shard d, slab s of `recv` comes from shard s)."""
    dist.all_to_all_single(recv, send)


def _check_stream(ctx, device):
    """The legacy helpers below run their collectives on torch's current stream; the library launches on the context's
    own.  Unless they are one and the same, kernels and NCCL traffic race."""
    if getattr(device, "type", str(device)) != "cuda":
        return
    cur = torch.cuda.current_stream(device).cuda_stream
    if (ctx.opts.stream or 0) != cur:
        raise RuntimeError("the context must launch on torch's current stream (Context(stream=torch.cuda.current_stream().cuda_stream))")


def stage1_scan_sharded(ctx, n_reads, rank, world, dist, device):
    """Scan-sharded stage 1 over a cabi.Context: scan my slice of the reads, all-to-all the items over NCCL, count.
    Returns this shard's share of edge_counting (numpy int64[65536])."""
    _check_stream(ctx, device)
    if getattr(ctx, "n_short", n_reads) < n_reads:    # assist reads: the replicated scan tells their occurrences apart
        return ctx.stage1()
    lo, hi = read_range(n_reads, rank, world)
    slab = ctx.stage1_slab_items()                    # the same on every shard: computed from start_idx, no collective
    for attempt in range(2):
        need = ctx.stage1_scan(lo, hi, slab)
        counts = ctx.stage1_exchange_buffers()[3] if need <= slab else [0] * world
        got, need = share_counts(rank, world, dist, device, counts, need)
        if need <= slab:
            break
        slab = need                                   # a send slab overflowed somewhere (skewed input): one rescan that fits
    else:
        raise RuntimeError("stage-1 send slabs overflow after the rescan")
    sp, rp, slab_bytes, _ = ctx.stage1_exchange_buffers()
    send = torch.as_tensor(DevBuf(sp, world * slab_bytes), device=device)
    recv = torch.as_tensor(DevBuf(rp, world * slab_bytes), device=device)
    exchange_items(rank, world, dist, device, send, recv)
    return ctx.stage1_count(got)


def exchange_ctx(ctx, rank, world, dist, device):
    """The same over a cabi.Context (device pointers from the C ABI)."""
    _check_stream(ctx, device)
    _, n, w = ctx.edges_local()

    def reserve(total, off):
        p = ctx.edges_reserve(total, off)
        return torch.as_tensor(DevBuf(p, max(total, 1) * w * 4), device=device)

    hp, hb = ctx.edge_hist_device_buffer()
    return exchange(rank, world, dist, device, n, w, reserve, torch.as_tensor(DevBuf(hp, hb), device=device))


# ---- the library-driven protocol (mgta_sharded_*): the caller only runs collectives ------------------------------------
def run_collective_tensors(op, send, recv, rank, world, dist):
    """One collective of the sharded build on tensors (uint8 views of the library's buffers; CPU tensors under gloo in
    the tests).  op: cabi.COLL_*.  ALL_GATHER: send is the slice recv[rank * n : (rank + 1) * n] (in place)."""
    from . import cabi
    if op == cabi.COLL_ALL_TO_ALL:
        dist.all_to_all_single(recv, send)
    elif op == cabi.COLL_ALL_GATHER:
        dist.all_gather_into_tensor(recv, send)
    elif op == cabi.COLL_ALL_REDUCE_SUM_U32:
        dist.all_reduce(recv.view(torch.int32), op=dist.ReduceOp.SUM)       # two's complement: the same bits as the u32 sum
    elif op == cabi.COLL_ALL_REDUCE_SUM_U64:
        dist.all_reduce(recv.view(torch.int64), op=dist.ReduceOp.SUM)
    else:
        raise ValueError("unknown collective %r" % (op,))


class ByteBuf:
    """uint8 view of raw device memory for torch.as_tensor (CUDA array interface)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}


class TorchComm:
    """Runs the library's collectives with torch.distributed (NCCL over NVLink).  The context must launch on torch's
    CURRENT stream: the library orders its kernels only with that stream, and torch orders the collective with the
    current stream too -- any other arrangement races (checked on every call)."""

    def __init__(self, ctx, rank, world, dist, device):
        self.ctx, self.rank, self.world, self.dist, self.device = ctx, rank, world, dist, device

    def __call__(self, c):
        from . import cabi
        cur = torch.cuda.current_stream(self.device).cuda_stream
        if (self.ctx.opts.stream or 0) != cur:
            raise RuntimeError("the context launches on stream %#x but torch's current stream is %#x: create the Context with "
                               "stream=torch.cuda.current_stream().cuda_stream inside `with torch.cuda.stream(...)`"
                               % (self.ctx.opts.stream or 0, cur))
        w, n = self.world, c.bytes
        if c.op == cabi.COLL_ALL_TO_ALL:
            send = torch.as_tensor(ByteBuf(c.send, w * n), device=self.device)
            recv = torch.as_tensor(ByteBuf(c.recv, w * n), device=self.device)
        elif c.op == cabi.COLL_ALL_GATHER:
            recv = torch.as_tensor(ByteBuf(c.recv, w * n), device=self.device)
            send = recv[self.rank * n:(self.rank + 1) * n]
        else:
            recv = torch.as_tensor(ByteBuf(c.recv, n), device=self.device)
            send = recv
        run_collective_tensors(c.op, send, recv, self.rank, w, self.dist)


def build_sharded(ctx, rank, world, dist, device, collect=True):
    """Both stages of a sharded build over a cabi.Context -> (edge_counting or None, stage-2 result of this shard)"""
    comm = TorchComm(ctx, rank, world, dist, device)
    ec = ctx.sharded(1, comm) if ctx.m > 1 else None
    return ec, ctx.sharded(2, comm, collect=collect)
