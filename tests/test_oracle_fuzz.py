"""CPU: the C restatement (oracle/cx1_oracle.c) against the UNMODIFIED reference binary RUN LIVE on seeded random read sets
(beyond the fixed goldens of tests/test_oracle.py): random genome size, coverage, ragged read lengths (some shorter than
k + 1), error rate, k from 9 to 127 around every word boundary, min-count 1..3, with and without mercy edges and assist
reads.  Compared: the bucket-ordered record stream, the per-bucket table, .counting, "Number mercy", the mercy candidates."""
import os
import re
import subprocess

import numpy as np
import pytest

from megagta_b200 import sdbg_io, synth
from oracle import oracle as O

KS = [9, 10, 15, 16, 17, 21, 31, 32, 33, 47, 48, 49, 63, 64, 65, 79, 80, 81, 95, 96, 97, 111, 112, 113, 126, 127]


def make_case(seed, d):
    rng = np.random.default_rng(1000 + seed)
    k = int(KS[seed % len(KS)])
    m = int(rng.integers(1, 4))
    mercy = bool(m > 1 and rng.random() < 0.5)
    assist = bool(rng.random() < 0.3)
    glen = int(rng.integers(3 * k + 80, 5 * k + 600))
    g = rng.integers(0, 4, glen, dtype=np.uint8)
    if rng.random() < 0.3:                                         # a tandem repeat and a reverse-complement palindrome
        g[10:10 + 2 * (k // 2 + 3)] = np.tile(g[10:12], k // 2 + 3)
        h = g[glen // 2:glen // 2 + k // 2 + 2].copy()
        g[glen // 2 + len(h):glen // 2 + 2 * len(h)] = (3 - h[::-1])[:len(g[glen // 2 + len(h):glen // 2 + 2 * len(h)])]
    cover = float(rng.choice([2, 6, 30]))
    lmin, lmax = k - 3, min(glen, k + int(rng.integers(4, 120)))
    err = float(rng.choice([0.0, 0.005, 0.03]))
    reads, bases = [], 0
    while bases < cover * glen:
        L = int(rng.integers(max(1, lmin), lmax + 1))
        p = int(rng.integers(0, glen - L + 1))
        r = g[p:p + L].copy()
        if rng.random() < 0.5:
            r = 3 - r[::-1]
        e = rng.random(L) < err
        r[e] = (r[e] + rng.integers(1, 4, int(e.sum()), dtype=np.uint8)) % 4
        reads.append(r)
        bases += L
    prefix = os.path.join(d, "f%d" % seed)
    synth.write_variable_reads(prefix, reads)
    fa = None
    if assist:
        fa = prefix + ".assist.fa"
        with open(fa, "w") as f:
            n, nb = int(rng.integers(1, 5)), 0
            for i in range(n):
                L = int(rng.integers(k + 1, min(glen, 3 * k + 50) + 1))
                p = int(rng.integers(0, glen - L + 1))
                s = "".join("ACGT"[c] for c in g[p:p + L])
                f.write(">a%d\n%s\n" % (i, s))
                nb += L
        open(fa + ".info", "w").write("%d %d\n" % (n, nb))
    return prefix, k, m, mercy, fa


@pytest.mark.parametrize("seed", range(52))
def test_oracle_equals_the_reference_on_random_inputs(seed, tmp_path):
    if not O.have_ref():
        pytest.skip("oracle/_ref/megagta_ref not built")
    d = str(tmp_path)
    prefix, k, m, mercy, fa = make_case(seed, d)
    out = os.path.join(d, "ref")
    try:
        log = O.run_ref_buildgraph(prefix, out, k, m, threads=2, need_mercy=mercy, assist_seq=fa)
    except subprocess.CalledProcessError:
        pytest.skip("the reference itself fails on this input (no item in any bucket: cx1.h aborts)")
    hdr, stream, meta = sdbg_io.canonical(out)
    rd = O.load_read_lib(prefix)
    n_short = None
    if fa:
        rd, n_short = O.with_assist(rd, fa)
    res = O.build_graph(rd, k, m, mercy, n_short)
    assert res["stream"] == stream, (k, m, mercy, bool(fa))
    assert np.array_equal(res["meta"], meta)
    assert hdr["total_size"] == int(res["meta"][:, 0].sum())
    if m > 1:
        assert O.counting_text(res["counting"]) == open(out + ".counting").read()
    if mercy:
        mm = re.search(r"Number mercy: (\d+)", log)
        assert mm and int(mm.group(1)) == int(res["num_mercy"])
        cands = np.concatenate([np.fromfile(os.path.join(d, x), dtype="<u8") for x in sorted(os.listdir(d))
                                if x.startswith("ref.mercy_cand.")] or [np.empty(0, "<u8")])
        mine = O.stage1(rd, k, m, True, n_short)[2]
        assert np.array_equal(np.sort(cands), np.sort(mine))


@pytest.mark.parametrize("seed", range(0, 52, 2))
def test_sdbg_oracle_equals_the_reference_loader_on_random_graphs(seed, tmp_path):
    """oracle/sdbg_oracle.py (numpy restatement of SuccinctDBG::LoadFromMultiFile + the rank / select builds) against dumps of
    the reference's own members (`megagta_ref sdbgdump`) for the graphs of the random read sets above, both multiplicity modes"""
    from oracle import sdbg_oracle as SO
    if not O.have_ref() or "sdbgdump" not in open(O.REF_BIN, "rb").read().decode("latin1"):
        pytest.skip("oracle/_ref/megagta_ref with sdbgdump not built")
    d = str(tmp_path)
    prefix, k, m, mercy, fa = make_case(seed, d)
    out = os.path.join(d, "ref")
    try:
        O.run_ref_buildgraph(prefix, out, k, m, threads=2, need_mercy=mercy, assist_seq=fa)
    except subprocess.CalledProcessError:
        pytest.skip("the reference itself fails on this input (no solid edge)")
    hdr, stream, meta = sdbg_io.canonical(out)
    for need_mult in (1, 0):
        ref = SO.ref_dump(O.REF_BIN, out, need_mult, os.path.join(d, "dump"))
        ref.pop("_load_seconds", None)
        want = SO.build(stream, np.asarray(meta), k, bool(need_mult))
        bad = [sec for sec in ref if want.get(sec) != ref[sec]]
        assert not bad, (bad, k, m)
