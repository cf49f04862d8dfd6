"""A/B timing of kernel variants at a chosen size (not a pytest file).
usage: gpu_ab.py s1 N_READS L K M ENV=V1,V2,...     stage-1 variants selected by an environment switch
       gpu_ab.py s2 N_READS [k]                      stage-2 sort variants"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from megagta_b200 import cabi, synth

S1_KEYS = ("ms_total", "ms_extract", "ms_partition", "ms_sort_emit", "n_items", "n_edges", "n_batches", "n_giants", "msd_levels")


def ab_s1(n, L, k, m, env, values):
    seq, start = synth.packed_metagenome(n, L, procs=16)
    with cabi.Context(k, m) as ctx:
        ctx.set_reads(seq, start, max_len=L)
        ref = None
        for v in values + values:
            if v == "-":
                os.environ.pop(env, None)
            else:
                os.environ[env] = v
            ec = ctx.stage1()
            if ref is None:
                ref = ec.copy()
            s1 = ctx.stats(1)
            print("k=%d %s=%s same=%s" % (k, env, v, bool((ec == ref).all())), {x: round(s1[x], 2) if isinstance(s1[x], float) else s1[x] for x in S1_KEYS}, flush=True)


def ab_s2(n, k):
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


if __name__ == "__main__":
    if sys.argv[1] == "s1":
        n, L, k, m = (int(x) for x in sys.argv[2:6])
        env, vals = sys.argv[6].split("=")
        ab_s1(n, L, k, m, env, vals.split(","))
    else:
        ab_s2(int(sys.argv[2]), int(sys.argv[3]) if len(sys.argv) > 3 else 31)
