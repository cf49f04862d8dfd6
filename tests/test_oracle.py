"""CPU tests: pin the C restatement (oracle/cx1_oracle.c) against golden values produced by the
unmodified reference binary (tests/golden/make_golden.py; SURVEY.md Appendix C lists the shared rows)."""
import hashlib

import numpy as np
import pytest

import oracle_memo as OM
from oracle import oracle as O

FAST_CASES = [
    "tiny_k21_m1", "tiny_k25_m2", "tiny_k25_m2_mercy",
    "smoke_k31_m2", "smoke_k31_m2_mercy", "smoke_k21_m2", "smoke_k22_m2", "smoke_k32_m2", "smoke_k41_m2",
    "smoke_k61_m2", "smoke_k99_m2", "smoke_k31_m1", "smoke_k31_m3", "smoke_k27_m3_mercy",
    "xander_k29_m1", "xander_k44_m1", "xander_k29_m2_mercy",
    "adversarial_k31_m2", "adversarial_k21_m1", "adversarial_k27_m3", "adversarial_k30_m2",
    "adversarial_k31_m2_mercy", "adversarial_k27_m3_mercy", "adversarial_k48_m2", "adversarial_k17_m2",
    "xander_k127_m1", "xander_k126_m2", "xander_k112_m2_mercy", "tiny_k10_m2", "tiny_k11_m1",      # the ends of the k range (kMaxK = 127)
]


def check_against_golden(res, g, mercy_cands=None):
    assert len(res["stream"]) == g["stream_bytes"]
    assert O.stream_hash(res["stream"]) == g["stream_hash"]
    assert O.meta_hash(res["meta"]) == g["meta_hash"]
    meta = np.asarray(res["meta"])
    assert int(meta[:, 0].sum()) == g["total_size"]
    assert int(meta[:, 1].sum()) == g["num_tips"]
    assert int(meta[:, 2].sum()) == g["large_multi"]
    if g.get("num_w") is not None and res.get("totals") is not None:
        assert [int(x) for x in res["totals"][:9]] == g["num_w"]
    if g["m"] > 1:
        txt = O.counting_text(res["counting"])
        assert hashlib.sha256(txt.encode()).hexdigest()[:16] == g["counting_sha"]
    if g["mercy"]:
        assert int(res["num_mercy"]) == g["num_mercy"]
        if mercy_cands is not None:
            assert len(mercy_cands) == g["mercy_cand_n"]
            assert hashlib.sha256(np.sort(mercy_cands).astype("<u8").tobytes()).hexdigest()[:16] == g["mercy_cand_sha"]


@pytest.mark.parametrize("case", FAST_CASES)
def test_oracle_matches_reference_golden(case, golden, read_lib):
    g = golden["cases"][case]
    _, rd = read_lib(g["dataset"])
    k, m, mercy = g["k"], g["m"], g["mercy"]
    solid = ec = cands = None                                       # O.build_graph in pieces: the pieces are shared (memo) with
    num_mercy = 0                                                   # the logic tests, which ask for the same graphs
    if m > 1:
        solid, ec, cands = OM.stage1(rd, k, m, mercy)
        if mercy:
            num_mercy = O.mercy(rd, k, solid, cands)
    stream, meta, totals = OM.stage2(rd, k, m, solid)
    res = dict(stream=stream, meta=meta, totals=totals, counting=ec, is_solid=solid, num_mercy=num_mercy)
    check_against_golden(res, g, cands if mercy else None)


def test_oracle_histograms_sum_to_item_counts(read_lib):
    _, rd = read_lib("smoke")
    k = 31
    h1 = O.s1_hist(rd, k)
    lens = np.diff(rd["start"].astype(np.int64))
    assert int(h1.sum()) == int(((lens - k + 4) * (lens >= k + 1)).sum())   # L-k+4 per read
    is_solid, _, _ = O.stage1(rd, k, 2)
    h2 = O.s2_hist(rd, k, 2, is_solid)
    _, meta, _ = O.stage2(rd, k, 2, is_solid)
    assert h2.sum() >= meta[:, 0].sum()
    assert np.all((h2 == 0) <= (meta[:, 0] == 0))


ASSIST_CASES = ["smoke_k31_m2_assist", "smoke_k31_m1_assist", "adversarial_k27_m3_assist"]


@pytest.mark.parametrize("case", ASSIST_CASES)
def test_oracle_with_assist_reads_matches_reference_golden(case, golden, read_lib, data_dir):
    """--assist_seq (reference s1.cpp:104-134): the FASTA sequences are appended reversed as always-solid reads that
    count in stage 1 but get no is_solid bits.  Golden = the unmodified reference binary run with --assist_seq."""
    import datasets
    g = golden["cases"][case]
    _, rd = read_lib(g["dataset"])
    fa = datasets.assist_fasta(g["dataset"], data_dir)
    assert datasets.md5(fa) == golden["datasets"][g["dataset"] + ".assist.fa"]
    rd2, n_short = O.with_assist(rd, fa)
    assert rd2["n_reads"] == n_short + 10
    res = O.build_graph(rd2, g["k"], g["m"], False, n_short=n_short)
    check_against_golden(res, g)
