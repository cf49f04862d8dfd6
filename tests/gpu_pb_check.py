"""Stage-2 tiling cross-check at scale (not a pytest file): the SdBG stream must not depend on the prefix-tile bits.
usage: gpu_pb_check.py N_READS K PB[,PB...] [N_GENOMES]"""
import hashlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from megagta_b200 import cabi, synth

n = int(sys.argv[1]); k = int(sys.argv[2]); pbs = sys.argv[3].split(",")
ng = int(sys.argv[4]) if len(sys.argv) > 4 else 64
seq, start = synth.packed_metagenome(n, 150, procs=16, n_genomes=ng)
ref = None
for pb in pbs:
    os.environ["MGTA_S2_PB"] = pb
    with cabi.Context(k, 2) as ctx:
        ctx.set_reads(seq, start, max_len=150)
        ec = ctx.stage1()
        stream, meta, totals = ctx.stage2()
        s2 = ctx.stats(2)
        sig = (hashlib.md5(stream).hexdigest(), int(meta[:, 0].sum()), int(meta[:, 1].sum()), int(meta[:, 2].sum()), [int(x) for x in totals])
        if ref is None:
            ref = (sig, meta.copy())
        same = sig == ref[0]
        print("PB", pb, "edges", sig[1], "md5", sig[0], "msd_levels", s2["msd_levels"], "giants", s2["n_giants"], "batches", s2["n_batches"], "SAME" if same else "DIFFERENT", flush=True)
        if not same:
            d = np.nonzero((meta != ref[1]).any(axis=1))[0]
            print("   first differing buckets", d[:10], "count", len(d), "rows", meta[d[:3]].tolist(), ref[1][d[:3]].tolist())
