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


# ------------------------------------------------------------------------------------------------
# Parallel in-memory generator for the bench workloads (same model as metagenome_reads, one RNG stream
# per 1M-read chunk so chunks can be produced by worker processes).

_GENOME_CACHE = {}


def _genomes(seed, n_genomes, glen, sigma):
    key = (seed, n_genomes, tuple(glen), sigma)
    if key not in _GENOME_CACHE:          # filled by the parent before the worker pool forks: workers inherit it
        _GENOME_CACHE.clear()
        _GENOME_CACHE[key] = _make_genomes(seed, n_genomes, glen, sigma)
    return _GENOME_CACHE[key]


def _make_genomes(seed, n_genomes, glen, sigma):
    rng = np.random.default_rng(seed)
    gl = rng.integers(glen[0], glen[1], size=n_genomes)
    G = [rng.integers(0, 4, size=int(n), dtype=np.uint8) for n in gl]
    ab = rng.lognormal(0, sigma, size=n_genomes) * gl
    return G, ab / ab.sum()


def _chunk_reads(args):
    seed, ci, n, L, n_genomes, glen, sigma, err = args
    G, ab = _genomes(seed, n_genomes, glen, sigma)
    rng = np.random.default_rng([seed, 1 + ci])
    gi = rng.choice(n_genomes, size=n, p=ab)
    out = np.empty((n, L), dtype=np.uint8)
    for g in np.unique(gi):
        idx = np.nonzero(gi == g)[0]
        pos = rng.integers(0, len(G[g]) - L, size=len(idx))
        out[idx] = G[g][pos[:, None] + np.arange(L)[None, :]]
    rcm = rng.random(n) < 0.5
    out[rcm] = 3 - out[rcm][:, ::-1]
    e = rng.random((n, L), dtype=np.float32) < err
    out[e] = (out[e] + rng.integers(1, 4, size=int(e.sum()), dtype=np.uint8)) % 4
    return out


def _pack_stream(bases_flat):
    """uint8[m] (m % 16 == 0) -> u32[m/16], MSB first."""
    b = bases_flat.reshape(-1, 16).astype(np.uint32)
    sh = (2 * (15 - np.arange(16))).astype(np.uint32)
    return (b << sh).sum(axis=1, dtype=np.uint64).astype(np.uint32)


def _chunk_packed(args):
    """-> (reversed, bit-contiguous packed words of the chunk, forward `.bin` records or None)"""
    want_bin = args[-1]
    reads = _chunk_reads(args[:-1])
    n, L = reads.shape
    rev = np.ascontiguousarray(reads[:, ::-1]).reshape(-1)
    pad = (-len(rev)) % 16
    if pad:
        rev = np.concatenate([rev, np.zeros(pad, dtype=np.uint8)])
    return _pack_stream(rev), (pack_forward(reads) if want_bin else None)


def packed_metagenome(n_reads, read_len, seed=20261017, n_genomes=64, glen=(200_000, 2_000_000), sigma=1.5, err=0.01,
                      procs=None, bin_prefix=None, bin_reads=0, chunk=1_000_000, first_read=0):
    """In-memory reads as the reference holds them (reversed, bit-contiguous): -> (seq u32[], start u64[n+1]).
    Optionally also writes the first `bin_reads` reads as a reference read library at `bin_prefix`
    (the bounded sample the CPU baseline runs on).  first_read (a multiple of `chunk`): generate reads
    [first_read, first_read + n_reads) of the same stream -- every chunk has its own RNG stream, so the slices the shards
    generate concatenate to exactly the read set one process would generate."""
    import multiprocessing as mp
    import os
    assert (chunk * read_len) % 16 == 0 and first_read % chunk == 0
    procs = procs or min(32, os.cpu_count() or 1)
    jobs = []
    for ci, s in enumerate(range(0, n_reads, chunk)):
        n = min(chunk, n_reads - s)
        jobs.append((seed, first_read // chunk + ci, n, read_len, n_genomes, glen, sigma, err, bool(bin_prefix) and s < bin_reads))
    total = n_reads * read_len
    seq = np.zeros(total // 16 + 1, dtype=np.uint32)
    binf = open(bin_prefix + ".bin", "wb") if bin_prefix else None
    written = 0
    _genomes(seed, n_genomes, glen, sigma)
    with mp.get_context("fork").Pool(min(procs, len(jobs))) as pool:
        for ci, (words, rec) in enumerate(pool.imap(_chunk_packed, jobs)):
            w0 = ci * chunk * read_len // 16
            seq[w0:w0 + len(words)] = words
            if rec is not None and written < bin_reads:
                take = min(len(rec), bin_reads - written)
                binf.write(rec[:take].tobytes())
                written += take
    if binf:
        binf.close()
        with open(bin_prefix + ".lib_info", "w") as f:
            f.write("%d %d\nsynthetic\n0 %d %d se\n" % (written * read_len, written, written - 1, read_len))
    start = (np.arange(n_reads + 1, dtype=np.uint64) * np.uint64(read_len))
    return seq, start
