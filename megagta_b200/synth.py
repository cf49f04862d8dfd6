"""Synthetic packed-read generators (the shapes BASELINE.json names).

Writes the reference's read-library files directly (format: sequence_manager.cpp:375-410 `.bin`,
read_lib_functions-inl.h:216-225 `.lib_info`): per read `u32 len` + ceil(len/16) u32 words, forward
orientation, first base in bits 31..30.
"""
import numpy as np


def pack_forward(bases_2d):
    """uint8[n, L] (values 0..3) -> '<u4'[n, 1+ceil(L/16)] records."""
    n, L = bases_2d.shape
    W = (L + 15) // 16
    pad = np.zeros((n, W * 16), dtype=np.uint32)
    pad[:, :L] = bases_2d
    sh = (2 * (15 - np.arange(16))).astype(np.uint32)
    words = (pad.reshape(n, W, 16) << sh).sum(axis=2, dtype=np.uint64).astype(np.uint32)
    rec = np.empty((n, W + 1), dtype="<u4")
    rec[:, 0] = L
    rec[:, 1:] = words
    return rec


def metagenome_reads(n_reads, read_len, seed=20261017, n_genomes=64, glen=(200_000, 2_000_000), sigma=1.5,
                     err=0.01, chunk=1_000_000):
    """Yield uint8[n, L] chunks of synthetic metagenome reads (SURVEY.md Appendix E.3 `gen_bin`)."""
    rng = np.random.default_rng(seed)
    gl = rng.integers(glen[0], glen[1], size=n_genomes)
    G = [rng.integers(0, 4, size=int(n), dtype=np.uint8) for n in gl]
    ab = rng.lognormal(0, sigma, size=n_genomes) * gl
    ab /= ab.sum()
    L = read_len
    for s in range(0, n_reads, chunk):
        n = min(chunk, n_reads - s)
        gi = rng.choice(n_genomes, size=n, p=ab)
        out = np.empty((n, L), dtype=np.uint8)
        for g in np.unique(gi):
            idx = np.nonzero(gi == g)[0]
            pos = rng.integers(0, len(G[g]) - L, size=len(idx))
            out[idx] = G[g][pos[:, None] + np.arange(L)[None, :]]
        rcm = rng.random(n) < 0.5
        out[rcm] = 3 - out[rcm][:, ::-1]
        e = rng.random((n, L)) < err
        out[e] = (out[e] + rng.integers(1, 4, size=int(e.sum()), dtype=np.uint8)) % 4
        yield out


def write_read_lib(prefix, chunks, read_len, n_reads):
    with open(prefix + ".bin", "wb") as f:
        for c in chunks:
            f.write(pack_forward(c).tobytes())
    with open(prefix + ".lib_info", "w") as f:
        f.write("%d %d\nsynthetic\n0 %d %d se\n" % (n_reads * read_len, n_reads, n_reads - 1, read_len))


def write_metagenome(prefix, n_reads, read_len, seed=20261017, **kw):
    write_read_lib(prefix, metagenome_reads(n_reads, read_len, seed, **kw), read_len, n_reads)


def write_variable_reads(prefix, reads):
    """reads: list of uint8 arrays (values 0..3) of arbitrary lengths."""
    total = 0
    with open(prefix + ".bin", "wb") as f:
        for r in reads:
            r = np.asarray(r, dtype=np.uint8)
            if len(r) == 0:
                f.write(np.array([0], dtype="<u4").tobytes())
                continue
            f.write(pack_forward(r[None, :]).tobytes())
            total += len(r)
    mx = max((len(r) for r in reads), default=0)
    with open(prefix + ".lib_info", "w") as f:
        f.write("%d %d\nsynthetic\n0 %d %d se\n" % (total, len(reads), len(reads) - 1, mx))
