"""Timing of the mercy pass at a chosen size (not a pytest file).  usage: gpu_mercy_time.py N_READS [k] [m]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from megagta_b200 import cabi, synth

n = int(sys.argv[1]); k = int(sys.argv[2]) if len(sys.argv) > 2 else 31; m = int(sys.argv[3]) if len(sys.argv) > 3 else 2
seq, start = synth.packed_metagenome(n, 150, procs=16)
for mercy in (False, True, True):
    with cabi.Context(k, m, need_mercy=mercy) as ctx:
        ctx.set_reads(seq, start, max_len=150)
        t = time.time(); ctx.stage1(); t1 = time.time() - t
        s1 = ctx.stats(1)
        t = time.time(); ctx.stage2(collect=False); t2 = time.time() - t
        s2 = ctx.stats(2)
        print("need_mercy", mercy, "stage1 wall %.3f s (device %.1f ms: extract %.1f partition %.1f count %.1f)" % (t1, s1["ms_total"], s1["ms_extract"], s1["ms_partition"], s1["ms_sort_emit"]),
              "stage2 wall %.3f s edges %d" % (t2, s2["n_edges"]), "num_mercy", ctx.num_mercy() if mercy else None,
              "candidates", len(ctx.mercy_candidates()) if mercy else None, flush=True)
