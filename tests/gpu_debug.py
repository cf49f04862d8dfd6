"""Ad-hoc GPU diagnostic (not a pytest file): one case, prints where the stream / table differ from the oracle."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import datasets
from megagta_b200 import cabi
from oracle import oracle as O

def run(ds, k, m, **kw):
    prefix = datasets.materialise(ds, "/tmp/mgta_data")
    rd = O.load_read_lib(prefix)
    exp_solid = None
    with cabi.Context(k, m, **kw) as ctx:
        ctx.set_reads(rd["seq"], rd["start"], max_len=rd["max_len"])
        if m > 1:
            ctx.stage1()
            exp_solid, _, _ = O.stage1(rd, k, m)
        stream, meta, totals = ctx.stage2()
        st = ctx.stats(2)
    es, em, et = O.stage2(rd, k, m, exp_solid)
    print(ds, k, m, kw, "stream", stream == es, len(stream), len(es), "meta", np.array_equal(meta, em), "totals", np.array_equal(totals, et), st["n_giants"], st["msd_levels"])
    if not np.array_equal(meta, em):
        bad = np.nonzero((meta != em).any(axis=1))[0]
        print("  bad buckets", len(bad), bad[:10], meta[bad[:5]].tolist(), em[bad[:5]].tolist())
    if stream != es:
        a = np.frombuffer(stream, np.uint8); b = np.frombuffer(es, np.uint8)
        n = min(len(a), len(b)); d = np.nonzero(a[:n] != b[:n])[0]
        print("  first diff byte", d[:5], "of", n)
        if len(d):
            p = int(d[0]) & ~1
            print("  got", a[max(0, p - 8):p + 16].tolist(), "exp", b[max(0, p - 8):p + 16].tolist())
    print("  totals", totals.tolist(), et.tolist(), flush=True)

if __name__ == "__main__" and len(sys.argv) == 1:
    run("tiny", 21, 1)
    run("tiny", 25, 2)
    run("smoke", 31, 2)
    run("smoke", 61, 2)

def dump(ds, k, m, nrec=24):
    from megagta_b200 import sdbg_io
    prefix = datasets.materialise(ds, "/tmp/mgta_data")
    rd = O.load_read_lib(prefix)
    with cabi.Context(k, m) as ctx:
        ctx.set_reads(rd["seq"], rd["start"], max_len=rd["max_len"])
        if m > 1:
            ctx.stage1()
        stream, meta, totals = ctx.stage2()
    es, em, et = O.stage2(rd, k, m, O.stage1(rd, k, m)[0] if m > 1 else None)
    wpt = (2 * k + 31) // 32
    g = sdbg_io.decode_stream(stream, wpt)
    e = sdbg_io.decode_stream(es, wpt)
    print(len(g), len(e))
    for i in range(nrec):
        print(i, g[i] if i < len(g) else None, e[i] if i < len(e) else None)

if len(sys.argv) > 1 and sys.argv[1] == "dump":
    dump("tiny", 21, 1)
