"""`megagta_b200 findstart` (model k-mers on the host, the scan of every read x strand x frame on the device:
mgta_find_seeds) against the UNMODIFIED reference `findstart` (oracle/_ref/megagta_ref; fast_kmer_filter.cpp:49-218) on a
synthetic gene family: an aligned reference with every special character the model-only k-mer generator treats differently
(lower-case inserts, '-', '.', 'X', 'x', '*'), reads shredded from back-translated members on both strands, random reads and
an extra contig file.  The seed lines are compared as sorted sets (the reference shuffles its output)."""
import hashlib
import json
import os
import random
import subprocess

import numpy as np
import pytest

import datasets
from megagta_b200 import synth
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "megagta_b200", "bin", "megagta_b200")
AA = "ARNDCQEGHILKMFPSTWYV"
CODE = "KNKNTTTTRSRSIIMIQHQHPPPPRRRRLLLLEDEDAAAAGGGGVVVV*Y*YSSSS*CWCLFLF"    # codon 16 b0 + 4 b1 + b2 -> residue, ACGT = 0123


def make_inputs(d, seed=5):
    rng = random.Random(seed)
    codons = {}
    for i, a in enumerate(CODE):
        codons.setdefault(a, []).append("ACGT"[i >> 4] + "ACGT"[(i >> 2) & 3] + "ACGT"[i & 3])
    anc = ["".join(rng.choice(AA) for _ in range(140)) for _ in range(3)]
    members, ref_lines = [], []
    for i in range(36):
        p = list(anc[i % 3])
        for j in range(len(p)):
            if rng.random() < 0.06:
                p[j] = rng.choice(AA)
        members.append("".join(p))
        row = ["." * rng.randrange(0, 9)]
        for j, c in enumerate(p):
            u = rng.random()
            if u < 0.03:
                row.append("-")                                # deleted model column
            elif u < 0.05:
                row.append(c + "".join(rng.choice(AA).lower() for _ in range(rng.randrange(1, 4))))   # insert states
            elif u < 0.06:
                row.append("X")
            elif u < 0.065:
                row.append("x")
            elif u < 0.07:
                row.append(c + "*")
            elif u < 0.09:
                row.append(c + "." * rng.randrange(1, 4))
            else:
                row.append(c)
        ref_lines.append(">ref%d some description  \n%s\n" % (i, "".join(row) + "." * rng.randrange(0, 7)))
    ref = os.path.join(d, "ref_aligned.faa")
    with open(ref, "w") as f:
        f.write("".join(ref_lines))
    comp = str.maketrans("ACGT", "TGCA")
    reads = []
    for m in members + anc:
        for _ in range(3):
            dna = "".join(rng.choice(codons[a]) for a in m)
            dna = "".join(rng.choice("ACGT") for _ in range(rng.randrange(0, 40))) + dna + "".join(rng.choice("ACGT") for _ in range(30))
            for _ in range(25):
                L = rng.randrange(30, 151)
                s = rng.randrange(0, max(1, len(dna) - L))
                r = dna[s:s + L]
                if rng.random() < 0.5:
                    r = r.translate(comp)[::-1]
                reads.append(r)
    reads += ["".join(rng.choice("ACGT") for _ in range(rng.randrange(10, 151))) for _ in range(2000)]
    rng.shuffle(reads)
    prefix = os.path.join(d, "reads")
    synth.write_variable_reads(prefix, [np.frombuffer(r.translate(str.maketrans("ACGT", "\0\1\2\3")).encode(), np.uint8) for r in reads])
    contigs = os.path.join(d, "contigs.fa")
    with open(contigs, "w") as f:
        for i, m in enumerate(members[:6]):
            f.write(">c%d\n%s\n" % (i, "".join(rng.choice(codons[a]) for a in m * 2)))
    return ref, prefix + ".bin", contigs


def run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    return sorted(r.stdout.splitlines())


def digest(lines):
    return hashlib.sha1("\n".join(lines).encode()).hexdigest()[:16]


CASES = [(45, False), (30, True), (60, False), (72, True), (44, False)]


@pytest.mark.gpu
@pytest.mark.parametrize("k_size,with_contigs", CASES)
def test_findstart_seeds_equal_the_reference(tmp_path, k_size, with_contigs):
    ref, binf, contigs = make_inputs(str(tmp_path))
    extra = ["2", contigs] if with_contigs else []
    ours = run([BIN, "findstart", ref, binf, str(k_size)] + extra)
    assert len(ours) > 100
    golden = json.load(open(os.path.join(datasets.GOLDEN_DIR, "findstart_golden.json")))
    assert digest(ours) == golden["k%d_contigs%d" % (k_size, int(with_contigs))]     # the reference's lines, made here (make_findstart_golden.py)
    if O.have_ref():
        theirs = run([O.REF_BIN, "findstart", ref, binf, str(k_size)] + extra)
        assert ours == theirs


def test_findstart_golden_inputs_are_reproducible(tmp_path):
    """CPU: the generator above still writes the inputs the goldens were made on"""
    ref, binf, contigs = make_inputs(str(tmp_path))
    golden = json.load(open(os.path.join(datasets.GOLDEN_DIR, "findstart_golden.json")))
    got = {n: hashlib.md5(open(p, "rb").read()).hexdigest() for n, p in (("ref", ref), ("bin", binf), ("contigs", contigs))}
    assert got == golden["inputs"]
