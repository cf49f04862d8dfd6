"""`megagta_b200 buildlib` (host kseq-rule parser + mgta_pack_reads on the device) against the UNMODIFIED reference
`buildlib` (oracle/_ref/megagta_ref; read_lib_functions-inl.h:119-226): byte-identical <P>.bin and <P>.lib_info for
single-end FASTA (multi-line, lower case, Ns, CRLF, empty lines), FASTQ (gzip'ed), paired and interleaved libraries, and
the reference's own in-tree fixture (tests/golden/xander.bin was packed by the reference from test_reads.fa)."""
import gzip
import os
import random
import subprocess

import numpy as np
import pytest

from megagta_b200 import cabi
from oracle import oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "megagta_b200", "bin", "megagta_b200")


def rand_seq(rng, n, alphabet="ACGT"):
    return "".join(rng.choice(alphabet) for _ in range(n))


def write_inputs(d):
    rng = random.Random(11)
    fa = os.path.join(d, "a.fa")
    with open(fa, "w", newline="") as f:
        for i in range(3000):
            s = rand_seq(rng, rng.randrange(0, 260), "ACGTNacgtn" if i % 7 == 0 else "ACGT")
            f.write(">r%d some comment\n" % i)
            eol = "\r\n" if i % 5 == 0 else "\n"
            for o in range(0, len(s), 60):
                f.write(s[o:o + 60] + eol)
            if i % 11 == 0:
                f.write("\n")
    fq = os.path.join(d, "b.fq.gz")
    with gzip.open(fq, "wt") as f:
        for i in range(4000):
            s = rand_seq(rng, rng.randrange(1, 151))
            f.write("@q%d\n%s\n+\n%s\n" % (i, s, "".join(chr(33 + rng.randrange(40)) for _ in s)))   # quality may start with '@' or '>'
    p1, p2 = os.path.join(d, "p_1.fq"), os.path.join(d, "p_2.fq")
    with open(p1, "w") as f1, open(p2, "w") as f2:
        for i in range(2500):
            for f in (f1, f2):
                s = rand_seq(rng, 100)
                f.write("@p%d/1\n%s\n+p%d\n%s\n" % (i, s, i, "I" * 100))
    il = os.path.join(d, "i.fa")
    with open(il, "w") as f:
        for i in range(2000):
            f.write(">i%d\n%s\n" % (i, rand_seq(rng, 16 * rng.randrange(1, 9))))    # lengths that are multiples of 16
    lib = os.path.join(d, "reads.lib")
    with open(lib, "w") as f:
        f.write("a.fa single end fasta\nse %s\n" % fa)
        f.write("b.fq.gz\nse %s\n" % fq)
        f.write("p_1.fq,p_2.fq\npe %s %s\n" % (p1, p2))
        f.write("interleaved lib\ninterleaved %s\n" % il)
    return lib


def test_buildlib_files_equal_the_reference(tmp_path):
    if not O.have_ref():
        pytest.skip("oracle/_ref/megagta_ref not built")
    lib = write_inputs(str(tmp_path))
    ours, ref = str(tmp_path / "ours"), str(tmp_path / "ref")
    r = subprocess.run([BIN, "buildlib", lib, ours], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    O.run_ref_buildlib(lib, ref)
    assert open(ours + ".lib_info").read() == open(ref + ".lib_info").read()
    a, b = open(ours + ".bin", "rb").read(), open(ref + ".bin", "rb").read()
    assert len(a) == len(b) and a == b


def test_pack_reads_matches_the_in_tree_fixture():
    """tests/golden/xander.bin = the reference's buildlib on its in-tree test_reads.fa: unpack it, pack it again on the device"""
    import datasets
    raw = np.fromfile(os.path.join(datasets.GOLDEN_DIR, "xander.bin"), dtype="<u4")
    seqs, p = [], 0
    while p < len(raw):
        n = int(raw[p]); w = raw[p + 1:p + 1 + (n + 15) // 16]
        codes = ((w[:, None] >> (30 - 2 * np.arange(16, dtype=np.uint32))) & 3).reshape(-1)[:n]
        seqs.append(bytes(np.frombuffer(b"ACGT", np.uint8)[codes]))
        p += 1 + (n + 15) // 16
    assert np.array_equal(cabi.pack_reads(seqs), raw)
    assert np.array_equal(cabi.pack_reads([b"", b"acgtn", b"N" * 33]), np.array([0, 5, 0x1B800000, 33, 0xAAAAAAAA, 0xAAAAAAAA, 0x80000000], np.uint32))
