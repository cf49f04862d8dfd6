"""Generate tests/golden/sdbg_golden.json: digests of the arrays the UNMODIFIED reference holds after
SuccinctDBG::LoadFromMultiFile on graphs its own buildgraph wrote (oracle/_ref/megagta_ref buildgraph + sdbgdump).
Run here (the container that has /root/reference):  python tests/golden/make_sdbg_golden.py"""
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import datasets  # noqa: E402
from oracle import oracle as O  # noqa: E402
from oracle import sdbg_oracle as SO  # noqa: E402

CASES = [("tiny", 25, 2), ("smoke", 31, 2), ("smoke", 21, 1), ("smoke", 61, 2), ("adversarial", 31, 2), ("adversarial", 21, 1),
         ("xander", 29, 1), ("meta200k", 31, 2)]


def main():
    O.build()
    assert O.have_ref()
    work = tempfile.mkdtemp(prefix="mgta_sdbg_golden_")
    golden = {}
    for ds, k, m in CASES:
        prefix = datasets.materialise(ds, os.path.join(work, "data"))
        out = os.path.join(work, "%s_k%d_m%d" % (ds, k, m))
        O.run_ref_buildgraph(prefix, out, k, m, threads=4)
        for need_mult in (1, 0):
            d = SO.ref_dump(O.REF_BIN, out, need_mult, out + ".dump")
            golden["%s_k%d_m%d_mult%d" % (ds, k, m, need_mult)] = SO.digest(d)
            print(ds, k, m, need_mult, len(d), "sections")
    json.dump(golden, open(os.path.join(datasets.GOLDEN_DIR, "sdbg_golden.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
