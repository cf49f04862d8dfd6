"""Multi-GPU parity driver (not a pytest file; tests/test_gpu_multi.py launches it under torchrun).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tests/gpu_multi.py [case ...]

One process per GPU over NCCL, exactly the call chain bench.py times: rank 0 uploads the packed reads, ncclBroadcast
to the other ranks, then mgta_sharded_begin / _step per stage with shards.TorchComm running the collectives the library
asks for (scan-sharded stage 1: items all-to-all by hash owner; stage 2: exchange of the solid edges, bucket-sharded emission).  The shard streams are concatenated in rank order (= bucket order) on rank 0 and compared with the goldens
the UNMODIFIED reference binary produced (tests/golden/golden.json): stream hash, per-bucket table hash, w totals and
the .counting text."""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import datasets  # noqa: E402
from megagta_b200 import cabi, shards  # noqa: E402
from oracle import oracle as O  # noqa: E402


def main():
    cases = sys.argv[1:] or ["meta200k_k31_m2", "meta200k_k61_m2", "adversarial_k27_m3", "smoke_k31_m1", "meta1m_k31_m2"]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    golden = json.load(open(os.path.join(datasets.GOLDEN_DIR, "golden.json")))
    data_dir = os.environ.get("MGTA_TEST_DATA") or "/tmp/mgta_multi_data"
    failures = []
    for case in cases:
        g = golden["cases"][case]
        k, m = g["k"], g["m"]
        shape = torch.zeros(4, dtype=torch.int64, device=dev)
        if rank == 0:
            os.makedirs(data_dir, exist_ok=True)
            prefix = datasets.materialise(g["dataset"], data_dir)
            rd = O.load_read_lib(prefix)                       # the oracle's loader: test-side file parsing only
            seq = np.ascontiguousarray(rd["seq"], dtype=np.uint32)
            start = np.ascontiguousarray(rd["start"], dtype=np.uint64)
            shape = torch.tensor([len(seq), len(start) - 1, int(start[-1]), rd["max_len"]], dtype=torch.int64, device=dev)
        dist.broadcast(shape, 0)
        n_words, n_reads, total_bases, max_len = [int(x) for x in shape.tolist()]
        stream = torch.cuda.Stream(device=dev)
        with torch.cuda.stream(stream), cabi.Context(k, m, device=local, rank=rank, world=world,
                                                     stream=stream.cuda_stream) as ctx:
            if rank == 0:
                ctx.set_reads(seq, start, max_len=max_len)
            else:
                ctx.alloc_reads(n_words, n_reads, n_reads, total_bases, max_len)
            (sp, sb), (tp, tb) = ctx.reads_device_buffers()
            dist.broadcast(torch.as_tensor(shards.ByteBuf(sp, sb), device=dev), 0)
            dist.broadcast(torch.as_tensor(shards.ByteBuf(tp, tb), device=dev), 0)
            # the library walks the protocol; TorchComm runs the collectives it asks for with NCCL
            ec_np, (st, meta, totals) = shards.build_sharded(ctx, rank, world, dist, dev)
            ec = torch.from_numpy(ec_np if ec_np is not None else np.zeros(65536, dtype=np.int64)).to(dev)
            lo, hi = ctx.shard_range()
        # gather on rank 0: streams in rank order, tables and totals summed
        meta_t = torch.from_numpy(meta).to(dev)
        tot_t = torch.from_numpy(totals).to(dev)
        dist.all_reduce(meta_t)
        dist.all_reduce(tot_t)
        sizes = torch.zeros(world, dtype=torch.int64, device=dev)
        sizes[rank] = len(st)
        dist.all_reduce(sizes)
        mine = torch.frombuffer(bytearray(st) if st else bytearray(1), dtype=torch.uint8).to(dev)[:len(st)]
        parts = []
        for r in range(world):
            buf = mine if r == rank else torch.empty(int(sizes[r]), dtype=torch.uint8, device=dev)
            if int(sizes[r]):
                dist.broadcast(buf, r)
            parts.append(buf)
        if rank == 0:
            whole = torch.cat(parts).cpu().numpy().tobytes()
            ok = (len(whole) == g["stream_bytes"] and O.stream_hash(whole) == g["stream_hash"]
                  and O.meta_hash(meta_t.cpu().numpy()) == g["meta_hash"]
                  and [int(x) for x in tot_t[:9].tolist()] == g["num_w"])
            if m > 1:
                txt = O.counting_text(ec.cpu().numpy())
                ok = ok and hashlib.sha256(txt.encode()).hexdigest()[:16] == g["counting_sha"]
            # edge_counting came back all-reduced: already whole on every rank
            print("multi-gpu parity %-22s world=%d shard bytes=%s : %s" % (case, world, [int(x) for x in sizes.tolist()],
                                                                           "OK" if ok else "MISMATCH"), flush=True)
            if not ok:
                failures.append(case)
        dist.barrier()
    bad = torch.tensor([len(failures)], device=dev)
    dist.broadcast(bad, 0)
    dist.destroy_process_group()
    sys.exit(1 if int(bad.item()) else 0)


if __name__ == "__main__":
    main()
