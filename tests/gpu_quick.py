"""Ad-hoc GPU diagnostic (not a pytest file): runs a few cases and prints where parity breaks."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import datasets
from megagta_b200 import cabi
from oracle import oracle as O

def run(ds, k, m, **kw):
    prefix = datasets.materialise(ds, "/tmp/mgta_data")
    rd = O.load_read_lib(prefix)
    t = time.time()
    with cabi.Context(k, m, **kw) as ctx:
        ctx.set_reads(rd["seq"], rd["start"], max_len=rd["max_len"])
        ok = True
        exp_solid = None
        if m > 1:
            h1 = ctx.histogram(1)
            e1 = O.s1_hist(rd, k)
            print("  s1 hist equal:", np.array_equal(h1, e1), int(h1.sum()), int(e1.sum()))
            ec = ctx.stage1()
            exp_solid, exp_ec, _ = O.stage1(rd, k, m)
            print("  counting equal:", np.array_equal(ec, exp_ec), ec[1:6], exp_ec[1:6], ctx.stats(1))
            got = ctx.get_is_solid(); n = O.solid_bytes(rd, k)
            print("  is_solid equal:", np.array_equal(got[:n], exp_solid[:n]), int(np.unpackbits(got[:n]).sum()), int(np.unpackbits(exp_solid[:n]).sum()))
        h2 = ctx.histogram(2)
        e2 = O.s2_hist(rd, k, m, exp_solid if exp_solid is not None else np.zeros(8, np.uint8))
        print("  s2 hist equal:", np.array_equal(h2, e2), int(h2.sum()), int(e2.sum()))
        stream, meta, totals = ctx.stage2()
        es, em, et = O.stage2(rd, k, m, exp_solid)
        print("  stream equal:", stream == es, len(stream), len(es), "meta", np.array_equal(meta, em), "totals", totals, et)
        print("  stats2", ctx.stats(2))
        if stream != es:
            a = np.frombuffer(stream, np.uint8); b = np.frombuffer(es, np.uint8)
            n = min(len(a), len(b)); d = np.nonzero(a[:n] != b[:n])[0]
            print("  first diff byte", d[:5] if len(d) else "prefix equal")
    print("%s k=%d m=%d %s: %.2fs" % (ds, k, m, kw, time.time() - t), flush=True)

if __name__ == "__main__":
    run("tiny", 21, 1)
    run("tiny", 25, 2)
    run("smoke", 31, 2)
    run("tiny", 25, 2, sort_items_cap=64)
    run("smoke", 31, 2, sort_items_cap=256)
    run("adversarial", 31, 2)
    run("smoke", 31, 2, hbm_budget_bytes=24 << 20)
