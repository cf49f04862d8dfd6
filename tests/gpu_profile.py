"""One buildgraph at a chosen size for ncu captures (not a pytest file).  usage: gpu_profile.py N_READS [k] [m]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from megagta_b200 import cabi, synth

n = int(sys.argv[1]); k = int(sys.argv[2]) if len(sys.argv) > 2 else 31; m = int(sys.argv[3]) if len(sys.argv) > 3 else 2
seq, start = synth.packed_metagenome(n, 150, procs=16)
with cabi.Context(k, m) as ctx:
    ctx.set_reads(seq, start, max_len=150)
    if m > 1:
        ctx.stage1()
        print(ctx.stats(1))
    nb, meta, tot = ctx.stage2(collect="count")
    print(ctx.stats(2))
