"""Generate tests/golden/findstart_golden.json: digests of the sorted seed lines the UNMODIFIED reference `findstart` prints
(oracle/_ref/megagta_ref) for the synthetic gene family of tests/test_gpu_findstart.py.  Run here:
python tests/golden/make_findstart_golden.py"""
import hashlib
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import datasets  # noqa: E402
import test_gpu_findstart as T  # noqa: E402
from oracle import oracle as O  # noqa: E402


def main():
    O.build()
    assert O.have_ref()
    d = tempfile.mkdtemp(prefix="mgta_findstart_golden_")
    ref, binf, contigs = T.make_inputs(d)
    golden = {"inputs": {n: hashlib.md5(open(p, "rb").read()).hexdigest() for n, p in (("ref", ref), ("bin", binf), ("contigs", contigs))}}
    for k_size, with_contigs in T.CASES:
        lines = T.run([O.REF_BIN, "findstart", ref, binf, str(k_size)] + (["2", contigs] if with_contigs else []))
        golden["k%d_contigs%d" % (k_size, int(with_contigs))] = T.digest(lines)
        print(k_size, with_contigs, len(lines), "seed lines")
    json.dump(golden, open(os.path.join(datasets.GOLDEN_DIR, "findstart_golden.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
