"""Multi-GPU build: the protocol lives in libmgta_cuda.so (mgta_sharded_begin / _step / _result, include/mgta_cuda.h); this
module only runs the collectives the library asks for with `torch.distributed` (NCCL over NVLink on the GPUs, gloo on
CPU tensors in the tests).  DESIGN.md section 7."""
import torch


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
