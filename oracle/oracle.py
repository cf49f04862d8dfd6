"""oracle/oracle.py -- TEST INFRASTRUCTURE ONLY (checker), never imported by the product package.

ctypes front-end for oracle/_build/libcx1_oracle.so (our C restatement of the reference's CX1
reads->SdBG path) and a runner for oracle/_ref/megagta_ref (the unmodified reference compiled by
oracle/Makefile).  Also holds an independent numpy loader of the reference's packed-read files so
the checker does not share input code with the product.

Reference anchors: input format sequence_manager.cpp:375-410 / read_lib_functions-inl.h:216-261,
reversed in-memory layout sequence_package.h:247-252,341-367.
"""
import ctypes
import hashlib
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "libcx1_oracle.so")
REF_BIN = os.path.join(HERE, "_ref", "megagta_ref")
NB = 65536


def build():
    """Compile the C restatement (and the reference binary when /root/reference exists)."""
    subprocess.run(["make", "-s", "-C", HERE, "all"], check=True)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.cx1o_mercy.restype = ctypes.c_int64
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


# ----------------------------------------------------------------------------- packed reads (numpy)
def load_bin_records(prefix):
    """Parse <prefix>.bin (per read: u32 len, ceil(len/16) u32 words, forward orientation) and
    line 1 of <prefix>.lib_info.  Returns list of (len, words) views lazily as (lens, offsets, raw)."""
    raw = np.fromfile(prefix + ".bin", dtype="<u4")
    with open(prefix + ".lib_info") as f:
        total_bases, num_reads = map(int, f.readline().split())
    lens = np.empty(num_reads, dtype=np.int64)
    offs = np.empty(num_reads, dtype=np.int64)
    pos = 0
    if num_reads and raw.size == num_reads * (1 + (int(raw[0]) + 15) // 16) and \
            np.all(raw[:: 1 + (int(raw[0]) + 15) // 16] == raw[0]):
        w = (int(raw[0]) + 15) // 16
        lens[:] = int(raw[0])
        offs[:] = np.arange(num_reads, dtype=np.int64) * (w + 1) + 1
    else:
        for i in range(num_reads):
            ln = int(raw[pos])
            lens[i] = ln
            offs[i] = pos + 1
            pos += 1 + (ln + 15) // 16
    assert int(lens.sum()) == total_bases
    return lens, offs, raw


def unpack_bases(lens, offs, raw):
    """-> list-free representation: (bases uint8 concatenated in FORWARD orientation, start idx)."""
    start = np.zeros(len(lens) + 1, dtype=np.uint64)
    np.cumsum(lens, out=start[1:].view(np.int64))
    total = int(start[-1])
    bases = np.empty(total, dtype=np.uint8)
    if len(lens) and np.all(lens == lens[0]):
        L = int(lens[0]); w = (L + 15) // 16
        words = raw[(offs[:, None] + np.arange(w)[None, :])]
        sh = (2 * (15 - np.arange(16))).astype(np.uint32)
        b = ((words[:, :, None] >> sh[None, None, :]) & 3).astype(np.uint8).reshape(len(lens), w * 16)
        bases[:] = b[:, :L].reshape(-1)
    else:
        for i in range(len(lens)):
            L = int(lens[i]); w = (L + 15) // 16
            words = raw[offs[i]: offs[i] + w]
            sh = (2 * (15 - np.arange(16))).astype(np.uint32)
            b = ((words[:, None] >> sh[None, :]) & 3).astype(np.uint8).reshape(-1)
            bases[int(start[i]): int(start[i]) + L] = b[:L]
    return bases, start


def pack_reversed(bases, start):
    """Reverse every read (not complemented) and pack bit-contiguously, MSB first, 16 bases/word."""
    n = len(start) - 1
    total = int(start[-1])
    rev = np.empty(total, dtype=np.uint8)
    lens = np.diff(start.astype(np.int64))
    if n and np.all(lens == lens[0]):
        L = int(lens[0])
        rev[:] = bases.reshape(n, L)[:, ::-1].reshape(-1)
    else:
        for i in range(n):
            s, e = int(start[i]), int(start[i + 1])
            rev[s:e] = bases[s:e][::-1]
    nw = total // 16 + 1
    pad = np.zeros(nw * 16, dtype=np.uint32)
    pad[:total] = rev
    sh = (2 * (15 - np.arange(16))).astype(np.uint32)
    seq = (pad.reshape(nw, 16) << sh[None, :]).sum(axis=1, dtype=np.uint64).astype(np.uint32)
    return np.ascontiguousarray(seq), np.ascontiguousarray(start.astype(np.uint64))


def load_read_lib(prefix):
    """-> dict(seq=u32[], start=u64[n+1], n_reads, max_len) as the reference holds reads in memory."""
    lens, offs, raw = load_bin_records(prefix)
    bases, start = unpack_bases(lens, offs, raw)
    seq, start = pack_reversed(bases, start)
    return dict(seq=seq, start=start, n_reads=len(lens), max_len=int(lens.max()) if len(lens) else 0)


def with_assist(rd, fasta):
    """rd + the sequences of a FASTA file appended as assist reads (reference s1.cpp:104-134 ->
    SequencePackage::AppendReverseSeq: reversed, un-trimmed, dna_map ACGTNacgtn -> 0123201232, sequence_package.h:67-69).
    -> (new read dict, n_short).  max_len stays that of the short reads (s1.cpp:119)."""
    code = np.zeros(256, dtype=np.uint8)
    for c, v in zip("ACGTNacgtn", "0123201232"):
        code[ord(c)] = int(v)
    seqs, cur = [], None
    for line in open(fasta):
        line = line.rstrip("\n")
        if line.startswith(">"):
            if cur is not None:
                seqs.append("".join(cur))
            cur = []
        elif cur is not None:
            cur.append(line)
    if cur is not None:
        seqs.append("".join(cur))
    lens, offs, raw = None, None, None
    n0 = rd["n_reads"]
    old_bases = unpack_stream(rd["seq"], int(rd["start"][-1]))
    extra = [code[np.frombuffer(q.encode(), dtype=np.uint8)][::-1] for q in seqs]
    bases = np.concatenate([old_bases] + extra) if extra else old_bases
    start = np.concatenate([rd["start"].astype(np.uint64),
                            (int(rd["start"][-1]) + np.cumsum([len(e) for e in extra])).astype(np.uint64)])
    pad = (-len(bases)) % 16
    b = np.concatenate([bases, np.zeros(pad + 16, dtype=np.uint8)]).reshape(-1, 16).astype(np.uint32)
    sh = (2 * (15 - np.arange(16))).astype(np.uint32)
    seq = (b << sh).sum(axis=1, dtype=np.uint64).astype(np.uint32)
    return dict(seq=seq, start=start, n_reads=n0 + len(seqs), max_len=rd["max_len"]), n0


def unpack_stream(seq, n_bases):
    """u32 words (16 bases each, MSB first) -> uint8 bases[n_bases]"""
    w = np.asarray(seq, dtype=np.uint32)[: (n_bases + 15) // 16]
    sh = (2 * (15 - np.arange(16))).astype(np.uint32)
    return ((w[:, None] >> sh[None, :]) & 3).astype(np.uint8).reshape(-1)[:n_bases]


# ----------------------------------------------------------------------------- oracle calls
def words_s1(k):
    return lib().cx1o_words_s1(k)


def words_s2(k):
    return lib().cx1o_words_s2(k)


def s1_hist(rd, k):
    h = np.zeros(NB, dtype=np.int64)
    lib().cx1o_s1_hist(_p(rd["seq"]), _p(rd["start"]), ctypes.c_int64(rd["n_reads"]), k, _p(h))
    return h


def s2_hist(rd, k, m, is_solid, n_short=None):
    h = np.zeros(NB, dtype=np.int64)
    n_short = rd["n_reads"] if n_short is None else n_short
    lib().cx1o_s2_hist(_p(rd["seq"]), _p(rd["start"]), ctypes.c_int64(rd["n_reads"]), ctypes.c_int64(n_short),
                       rd["max_len"], k, m, _p(is_solid), _p(h))
    return h


def solid_bytes(rd, k, n_short=None):
    n_short = rd["n_reads"] if n_short is None else n_short
    return (max(0, (rd["max_len"] - k)) * n_short + 7) // 8


def stage1(rd, k, m, need_mercy=False, n_short=None):
    """-> (is_solid u8[], edge_counting i64[65536], mercy candidates u64[] sorted)"""
    n_short = rd["n_reads"] if n_short is None else n_short
    is_solid = np.zeros(solid_bytes(rd, k, n_short) + 8, dtype=np.uint8)
    ec = np.zeros(NB, dtype=np.int64)
    mp = ctypes.c_void_p()
    mn = ctypes.c_int64(0)
    rc = lib().cx1o_stage1(_p(rd["seq"]), _p(rd["start"]), ctypes.c_int64(rd["n_reads"]), ctypes.c_int64(n_short),
                           rd["max_len"], k, m, _p(is_solid), _p(ec), int(need_mercy),
                           ctypes.byref(mp), ctypes.byref(mn))
    assert rc == 0
    cand = np.empty(mn.value, dtype=np.uint64)
    if mn.value:
        ctypes.memmove(_p(cand), mp, mn.value * 8)
    lib().cx1o_free(mp)
    return is_solid, ec, cand


def mercy(rd, k, is_solid, cand, n_short=None):
    n_short = rd["n_reads"] if n_short is None else n_short
    return lib().cx1o_mercy(_p(rd["seq"]), _p(rd["start"]), ctypes.c_int64(rd["n_reads"]), ctypes.c_int64(n_short),
                            rd["max_len"], k, _p(is_solid), _p(cand), ctypes.c_int64(len(cand)))


def stage2(rd, k, m, is_solid, n_short=None):
    """-> (stream bytes, meta i64[65536,3], totals i64[10])"""
    n_short = rd["n_reads"] if n_short is None else n_short
    if is_solid is None:
        is_solid = np.zeros(8, dtype=np.uint8)
    sp = ctypes.c_void_p()
    sn = ctypes.c_int64(0)
    meta = np.zeros((NB, 3), dtype=np.int64)
    totals = np.zeros(10, dtype=np.int64)
    rc = lib().cx1o_stage2(_p(rd["seq"]), _p(rd["start"]), ctypes.c_int64(rd["n_reads"]), ctypes.c_int64(n_short),
                           rd["max_len"], k, m, _p(is_solid), ctypes.byref(sp), ctypes.byref(sn), _p(meta), _p(totals))
    assert rc == 0
    stream = ctypes.string_at(sp, sn.value) if sn.value else b""
    lib().cx1o_free(sp)
    return stream, meta, totals


def build_graph(rd, k, m, need_mercy=False, n_short=None):
    """Whole path.  -> dict(stream, meta, totals, counting (None if m==1), is_solid, num_mercy)"""
    is_solid = None
    ec = None
    num_mercy = 0
    if m > 1:
        is_solid, ec, cand = stage1(rd, k, m, need_mercy, n_short)
        if need_mercy:
            num_mercy = mercy(rd, k, is_solid, cand, n_short)
    stream, meta, totals = stage2(rd, k, m, is_solid, n_short)
    return dict(stream=stream, meta=meta, totals=totals, counting=ec, is_solid=is_solid, num_mercy=num_mercy)


# ----------------------------------------------------------------------------- hashing
def stream_hash(stream):
    return hashlib.sha256(stream).hexdigest()[:16]


def meta_hash(meta):
    h = hashlib.sha256()
    m = np.asarray(meta)
    lines = "".join("%d %d %d %d\n" % (b, m[b, 0], m[b, 1], m[b, 2]) for b in range(NB))
    h.update(lines.encode())
    return h.hexdigest()[:16]


def counting_text(ec):
    """<prefix>.counting as the reference writes it (s1.cpp:923-930)."""
    acc = np.cumsum(np.asarray(ec, dtype=np.int64)[1:])
    return "".join("%d %d\n" % (i + 1, acc[i]) for i in range(NB - 1))


# ----------------------------------------------------------------------------- reference binary
def have_ref():
    return os.path.exists(REF_BIN)


def run_ref_buildgraph(read_lib_prefix, out_prefix, k, m, threads=None, need_mercy=False, host_mem=None,
                       assist_seq=None, capture=True):
    """Run the unmodified reference `buildgraph` (build_graph.cpp:33-135).  Returns stderr text."""
    threads = threads or max(2, os.cpu_count() or 2)
    if host_mem is None:
        host_mem = int(0.9 * os.sysconf("SC_PAGE_SIZE") * os.sysconf("SC_PHYS_PAGES"))
    cmd = [REF_BIN, "buildgraph", "-k", str(k), "-m", str(m), "--host_mem", str(host_mem), "--mem_flag", "1",
           "--num_cpu_threads", str(threads), "--num_output_threads", str(max(1, threads // 3)),
           "--read_lib_file", read_lib_prefix, "--output_prefix", out_prefix]
    if need_mercy:
        cmd.append("--need_mercy")
    if assist_seq:
        cmd += ["--assist_seq", assist_seq]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE if capture else None, check=True)
    return r.stderr.decode() if capture else ""


def run_ref_buildlib(lib_file, out_prefix):
    subprocess.run([REF_BIN, "buildlib", lib_file, out_prefix], check=True, stdout=subprocess.DEVNULL,
                   stderr=subprocess.DEVNULL)
