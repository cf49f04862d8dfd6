"""Phase timing of the multi-GPU step (not a pytest file): torchrun ... tests/gpu_multi_profile.py [reads_per_gpu] [k]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
from megagta_b200 import cabi, shards, synth

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n_per = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 31
L = 150; n_reads = n_per * world; n_words = n_reads * L // 16 + 1
if rank == 0:
    seq, start = synth.packed_metagenome(n_reads, L, procs=16)
    seq_pin = torch.from_numpy(seq).pin_memory(); start_pin = torch.from_numpy(start.view(np.int64)).pin_memory()
stream = torch.cuda.Stream(device=dev)
ctx = cabi.Context(k, 2, device=local, rank=rank, world=world, stream=stream.cuda_stream)
T = {}
def tick(name, t0):
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    T.setdefault(name, []).append((time.time() - t0) * 1000); return time.time()
with torch.cuda.stream(stream):
    for it in range(4):
        torch.cuda.synchronize(); dist.barrier(); t = time.time()
        if rank == 0:
            ctx._check(ctx.lib.mgta_set_reads(ctx.h, seq_pin.data_ptr(), n_words, start_pin.data_ptr(), n_reads, n_reads, L), "set_reads")
            ctx.n_short, ctx.max_len = n_reads, L
        else:
            ctx.alloc_reads(n_words, n_reads, n_reads, n_reads * L, L)
        t = tick("h2d", t)
        (sp, sb), (tp, tb) = ctx.reads_device_buffers()
        dist.broadcast(torch.as_tensor(shards.DevBuf(sp, sb), device=dev), 0)
        dist.broadcast(torch.as_tensor(shards.DevBuf(tp, tb), device=dev), 0)
        t = tick("bcast", t)
        lo, hi = shards.read_range(n_reads, rank, world)
        slab = ctx.stage1_slab_items()
        t = tick("scan_size", t)
        need = ctx.stage1_scan(lo, hi, slab)
        t = tick("scan", t)
        sp, rp, slab_bytes, counts = ctx.stage1_exchange_buffers()
        got, need = shards.share_counts(rank, world, dist, dev, counts, need)
        send = torch.as_tensor(shards.DevBuf(sp, world * slab_bytes), device=dev)
        recv = torch.as_tensor(shards.DevBuf(rp, world * slab_bytes), device=dev)
        shards.exchange_items(rank, world, dist, dev, send, recv)
        t = tick("a2a", t)
        ctx.stage1_count(got)
        t = tick("count", t)
        shards.exchange_ctx(ctx, rank, world, dist, dev)
        t = tick("edge_xchg", t)
        nbytes, meta, totals = ctx.stage2(collect="count")
        t = tick("stage2_d2h", t)
        ctx.stage2(collect=False)
        t = tick("stage2_dev", t)
def step():
    shards.stage1_scan_sharded(ctx, n_reads, rank, world, dist, dev)
    shards.exchange_ctx(ctx, rank, world, dist, dev)
    ctx.stage2(collect=False)
with torch.cuda.stream(stream):
    for rep in range(2):
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time(); e0.record(stream)
        for _ in range(3):
            step()
        e1.record(stream); torch.cuda.synchronize()
        if rank == 0:
            print("bench-like loop: %.1f ms/step (events), %.1f ms/step (wall)" % (e0.elapsed_time(e1) / 3, (time.time() - t0) * 1000 / 3), flush=True)
if rank == 0:
    print("world", world, "reads/gpu", n_per, "k", k, "slab_bytes", slab_bytes, {k_: [round(x, 1) for x in v[1:]] for k_, v in T.items()}, ctx.stats(1), flush=True)
ctx.close(); dist.barrier(); dist.destroy_process_group()
