"""Deterministic test inputs (test infrastructure).  Every dataset is written as a reference read
library (<prefix>.bin + <prefix>.lib_info); golden.json records the md5 of each .bin so a drifted RNG
stream fails loudly instead of silently changing the inputs."""
import hashlib
import os
import random

import numpy as np

from megagta_b200 import synth

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN_DIR = os.path.join(HERE, "golden")
CODE = {c: i for i, c in enumerate("ACGT")}


def _smoke(prefix):
    """SURVEY.md Appendix E.2 (CPython `random`, seed 1): 20k x 100 bp from one 20 kb genome."""
    random.seed(1)
    G = "".join(random.choice("ACGT") for _ in range(20000))
    comp = {"A": "T", "C": "G", "G": "C", "T": "A"}
    reads = np.empty((20000, 100), dtype=np.uint8)
    for i in range(20000):
        p = random.randrange(0, len(G) - 100)
        s = G[p:p + 100]
        if random.random() < 0.5:
            s = "".join(comp[c] for c in reversed(s))
        s = list(s)
        for j in range(100):
            if random.random() < 0.01:
                s[j] = random.choice("ACGT")
        reads[i] = [CODE[c] for c in s]
    synth.write_read_lib(prefix, [reads], 100, 20000)


def _adversarial(prefix):
    """Variable-length reads with palindromes, homopolymers, dinucleotide repeats, reads shorter than
    k+1, very high coverage (multiplicity > 254 and > 65535 items in one group) and empty-ish reads."""
    rng = np.random.default_rng(7)
    g = rng.integers(0, 4, size=3000, dtype=np.uint8)
    pal_half = rng.integers(0, 4, size=40, dtype=np.uint8)
    pal = np.concatenate([pal_half, 3 - pal_half[::-1]])          # reverse-complement palindrome, 80 bp
    g[1000:1080] = pal
    reads = []
    for i in range(30000):                                         # ~1000x coverage of a 3 kb genome
        L = int(rng.integers(20, 121))
        p = int(rng.integers(0, len(g) - L))
        r = g[p:p + L].copy()
        if rng.random() < 0.5:
            r = 3 - r[::-1]
        e = rng.random(L) < 0.005
        r[e] = (r[e] + rng.integers(1, 4, size=int(e.sum()), dtype=np.uint8)) % 4
        reads.append(r)
    for i in range(900):                                           # poly-A / poly-T: one giant group
        reads.append(np.full(int(rng.integers(90, 121)), 0 if i % 3 else 3, dtype=np.uint8))
    for i in range(300):                                           # (AC)n and (ACG)n repeats
        L = int(rng.integers(60, 121))
        unit = [0, 1] if i % 2 else [0, 1, 2]
        reads.append(np.array((unit * L)[:L], dtype=np.uint8))
    for i in range(50):                                            # exact palindromic reads
        reads.append(pal.copy())
    for L in (1, 5, 16, 17, 31, 32, 33):                           # shorter than any k+1 we test
        reads.append(rng.integers(0, 4, size=L, dtype=np.uint8))
    order = rng.permutation(len(reads))
    synth.write_variable_reads(prefix, [reads[i] for i in order])


def _meta(n, L):
    def f(prefix):
        synth.write_metagenome(prefix, n, L, seed=20261017, n_genomes=16, glen=(20_000, 200_000))
    return f


def _tiny(prefix):
    rng = np.random.default_rng(3)
    g = rng.integers(0, 4, size=400, dtype=np.uint8)
    reads = [g[p:p + 60].copy() for p in rng.integers(0, 340, size=200)]
    synth.write_variable_reads(prefix, reads)


DATASETS = {
    "tiny": _tiny,                 # 200 x 60 bp
    "smoke": _smoke,               # 20k x 100 bp
    "adversarial": _adversarial,   # ~31k variable-length reads
    "meta200k": _meta(200_000, 100),
    "meta1m": _meta(1_000_000, 100),   # BASELINE config 0 shape (1M x 100 bp), GPU tests only
}


def assist_fasta(name, directory):
    """Deterministic `--assist_seq` input for dataset `name` (reference s1.cpp:104-134: FASTA + `<file>.info` holding
    `num_seq num_bases`): contig-like stretches stitched from the dataset's own reads (forward and reverse complement),
    with lower case, N, a multi-line layout and one sequence shorter than any k+1.  -> path of the FASTA file."""
    from oracle import oracle as O
    os.makedirs(directory, exist_ok=True)
    path = os.path.join(directory, name + ".assist.fa")
    if os.path.exists(path) and os.path.exists(path + ".info"):
        return path
    prefix = materialise(name, directory)
    lens, offs, raw = O.load_bin_records(prefix)
    bases, start = O.unpack_bases(lens, offs, raw)
    rng = np.random.default_rng(77)
    seqs = []
    n = len(lens)
    for i in range(9):
        parts = []
        for r in rng.integers(0, n, size=int(rng.integers(2, 7))):
            parts.append(bases[int(start[r]):int(start[r + 1])])
        a = np.concatenate(parts) if parts else np.zeros(0, np.uint8)
        if i % 3 == 1:
            a = (3 - a)[::-1]
        t = list("ACGT"[int(c)] for c in a)
        for j in rng.integers(0, max(1, len(t)), size=3):
            if len(t):
                t[int(j)] = "N" if i % 2 else t[int(j)].lower()
        seqs.append("".join(t))
    seqs.append("ACGTNacgtTTGACCA")                      # 16 bases: shorter than k + 1
    with open(path, "w") as f:
        for i, q in enumerate(seqs):
            f.write(">assist_%d some description\n" % i)
            for o in range(0, len(q), 70):
                f.write(q[o:o + 70] + "\n")
    with open(path + ".info", "w") as f:
        f.write("%d %d\n" % (len(seqs), sum(len(q) for q in seqs)))
    return path


def md5(path):
    h = hashlib.md5()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 20), b""):
            h.update(blk)
    return h.hexdigest()


def materialise(name, directory):
    """Write dataset `name` under `directory` (cached) and return its read-library prefix."""
    os.makedirs(directory, exist_ok=True)
    prefix = os.path.join(directory, name)
    if not (os.path.exists(prefix + ".bin") and os.path.exists(prefix + ".lib_info")):
        if name == "xander":
            import shutil
            shutil.copy(os.path.join(GOLDEN_DIR, "xander.bin"), prefix + ".bin")
            shutil.copy(os.path.join(GOLDEN_DIR, "xander.lib_info"), prefix + ".lib_info")
        else:
            DATASETS[name](prefix)
    return prefix


def metagenome_in_memory(n_reads, read_len, seed=20261017, **kw):
    """The reads `synth.write_metagenome(prefix, n_reads, read_len, seed)` would write (SURVEY Appendix E.3 `gen_bin`), without
    the file: -> (seq u32[] reversed + bit-contiguous as the reference holds them, start u64[n+1], md5 of the `.bin` bytes).
    Chunks of 1M reads must end on a word boundary (1M * read_len % 16 == 0)."""
    assert (1_000_000 * read_len) % 16 == 0
    h = hashlib.md5()
    seq = np.zeros(n_reads * read_len // 16 + 1, dtype=np.uint32)
    at = 0
    for c in synth.metagenome_reads(n_reads, read_len, seed, **kw):
        h.update(synth.pack_forward(c).tobytes())
        rev = np.ascontiguousarray(c[:, ::-1]).reshape(-1)
        pad = (-len(rev)) % 16
        if pad:
            rev = np.concatenate([rev, np.zeros(pad, dtype=np.uint8)])
        w = synth._pack_stream(rev)
        seq[at:at + len(w)] = w
        at += len(c) * read_len // 16
    start = np.arange(n_reads + 1, dtype=np.uint64) * np.uint64(read_len)
    return seq, start, h.hexdigest()
