"""A/B timing of stage-2 sort variants at a chosen size (not a pytest file).  usage: gpu_ab.py N_READS [k]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from megagta_b200 import cabi, synth

n = int(sys.argv[1]); k = int(sys.argv[2]) if len(sys.argv) > 2 else 31
seq, start = synth.packed_metagenome(n, 150, procs=16)
with cabi.Context(k, 2) as ctx:
    ctx.set_reads(seq, start, max_len=150)
    ctx.stage1()
    s1 = ctx.stats(1)
    print("s1", {x: s1[x] for x in ("ms_total", "ms_extract", "ms_partition", "ms_sort_emit", "n_items", "n_edges")})
    for bits, big in (("0", "64"), ("12", "64"), ("12", "128"), ("12", "256"), ("12", "1024"), ("12", "5000")):
        os.environ["MGTA_SORT_BIN_BITS"] = bits
        os.environ["MGTA_SORT_BIG_BIN"] = big
        ctx.stage2(collect=False)
        s2 = ctx.stats(2)
        print("bin_bits", bits, big, {x: s2[x] for x in ("ms_total", "ms_extract", "ms_partition", "ms_sort_emit", "n_giants", "n_items", "n_edges")})
