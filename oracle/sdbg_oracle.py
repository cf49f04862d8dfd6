"""oracle/sdbg_oracle.py -- TEST INFRASTRUCTURE ONLY (checker), never imported by the product package.

numpy restatement of the reference's in-memory SdBG load + rank/select build:
SuccinctDBG::LoadFromMultiFile (succinct_dbg.cpp:595-723), SuccinctDBG::init (succinct_dbg.h:62-83),
RankAndSelect4Bits::Build (rank_and_select.h:81-150), RankAndSelect1Bit::Build (rank_and_select.h:420-487),
SdbgReader::read_info's f_ (sdbg_multi_io.h:253-268).  Pinned on dumps of the reference's own members
(`oracle/_ref/megagta_ref sdbgdump`, oracle/ref_main.cpp; tests/golden/sdbg_golden.json).
"""
import hashlib
import struct
import subprocess

import numpy as np

NB = 65536


def read_dump(path):
    """sections of an `sdbgdump` file -> {name: bytes}"""
    d, b, o = {}, open(path, "rb").read(), 0
    while o < len(b):
        nm = b[o:o + 24].split(b"\0")[0].decode()
        n = struct.unpack_from("<Q", b, o + 24)[0]
        d[nm] = b[o + 32:o + 32 + n]
        o += 32 + n
    return d


def ref_dump(ref_bin, prefix, need_mult, out):
    r = subprocess.run([ref_bin, "sdbgdump", prefix, "1" if need_mult else "0", out], check=True, capture_output=True, text=True)
    d = read_dump(out)
    for line in r.stderr.splitlines():
        if line.startswith("load_seconds"):
            d["_load_seconds"] = line.split()[1].encode()       # wall time of SuccinctDBG::LoadFromMultiFile alone (not a section)
    return d


def _pack_bits(bits):
    """bool[n] -> u64 words, bit i at word i / 64, position i % 64"""
    n = len(bits)
    pad = np.zeros((-n) % 64, dtype=np.uint8)
    return np.packbits(np.concatenate([bits.astype(np.uint8), pad]), bitorder="little").view("<u8")


def _rank_tables(occ_cum, n):
    """occ_cum[p] = occurrences in positions [0, p), p = 0..n -> (major i64, minor u16) exclusive counts at interval starts
    clamped to n (rank_and_select.h:103-126 / :450-464)"""
    n_minor, n_major = (n + 255) // 256 + 1, (n + 65535) // 65536 + 1
    at = lambda step, cnt: occ_cum[np.minimum(np.arange(cnt, dtype=np.int64) * step, n)]
    major = at(65536, n_major).astype(np.int64)
    occ = at(256, n_minor).astype(np.int64)
    minor = (occ - major[np.arange(n_minor) // 256]).astype(np.uint16)
    return major, minor, occ


def _select_table(occ, count):
    """rank_to_interval (rank_and_select.h:131-147): entry s = (first interval i with Occ(i) > 256 s) - 1; closing entry"""
    n_table = (count + 255) // 256 + 1
    t = np.empty(n_table, dtype=np.uint32)
    s = np.arange(n_table - 1, dtype=np.int64) * 256
    t[:n_table - 1] = np.searchsorted(occ, s, side="right") - 1
    t[n_table - 1] = len(occ) - 1
    return t


def build(stream, meta, k, need_mult=True):
    """stream: bucket-ordered record bytes; meta: int64[65536, 3] -> {section name: bytes} with the names of `sdbgdump`"""
    wpt = (2 * k + 31) // 32
    u = np.frombuffer(stream, dtype="<u2")
    size, n_tips = int(meta[:, 0].sum()), int(meta[:, 1].sum())
    w = np.zeros(size, np.uint8); last = np.zeros(size, bool); tip = np.zeros(size, bool); mul = np.zeros(size, np.uint8)
    tip_seq = np.zeros(n_tips * wpt, dtype="<u4")
    large = []
    p, t = 0, 0
    for i in range(size):                                          # the loop of succinct_dbg.cpp:651-711
        item = int(u[p]); p += 1
        w[i] = item & 15; last[i] = (item >> 4) & 1; tip[i] = (item >> 5) & 1; mul[i] = item >> 8
        if (item >> 8) == 255:
            large.append((i, int(u[p]))); p += 1
        if (item >> 5) & 1:
            lab = u[p:p + 2 * wpt].astype(np.uint32); p += 2 * wpt
            tip_seq[t * wpt:(t + 1) * wpt] = lab[0::2] | (lab[1::2] << 16); t += 1
    assert p == len(u) and t == n_tips
    out = {"hdr": np.array([size, k, n_tips, wpt], np.int64).tobytes()}
    acc = np.cumsum(meta[:, 0])
    f = np.array([-1, 0] + [int(acc[(q + 1) * (NB // 4) - 1]) for q in range(4)], np.int64)
    out["f"] = f.tobytes()
    wn = np.concatenate([w, np.zeros((-size) % 16, np.uint8)]).astype(np.uint64).reshape(-1, 16)
    out["w"] = (wn << (4 * np.arange(16, dtype=np.uint64))).sum(axis=1, dtype=np.uint64).astype("<u8").tobytes()
    out["last"] = _pack_bits(last).tobytes(); out["is_tip"] = _pack_bits(tip).tobytes()
    out["invalid"] = _pack_bits(tip | (w == 0)).tobytes()
    out["tip_seq"] = tip_seq.tobytes()
    if need_mult:
        out["edge_multi"] = mul.tobytes()
        out["large_multi"] = np.array(large, np.int64).reshape(-1, 2).tobytes()
    else:
        out["is_multi_1"] = _pack_bits(mul <= 1).tobytes()
    freq = np.zeros(9, np.int64)
    for c in range(9):
        cum = np.concatenate([[0], np.cumsum(w == c)])
        if c == 0:
            cum[-1] += (-size) % 16                                # CountCharInWord_(0, last word) also counts its zero padding
        major, minor, occ = _rank_tables(cum, size)
        freq[c] = cum[-1]
        out["w_major_%d" % c] = major.tobytes(); out["w_minor_%d" % c] = minor.tobytes()
        out["w_sel_%d" % c] = _select_table(occ, int(cum[-1])).tobytes()
    out["w_freq"] = freq.tobytes()
    cum = np.concatenate([[0], np.cumsum(last)])
    major, minor, occ = _rank_tables(cum, size)
    out["last_ones"] = np.int64(cum[-1]).tobytes(); out["last_major"] = major.tobytes(); out["last_minor"] = minor.tobytes()
    out["last_sel"] = _select_table(occ, int(cum[-1])).tobytes()
    # rank_f[i] = rs_last_.Rank(f[i] - 1): ones of last in [0, f[i]); f[0] = -1 reads before the array in the reference and gives 0
    out["rank_f"] = np.array([int(cum[min(max(int(x), 0), size)]) for x in f], np.int64).tobytes()
    cum = np.concatenate([[0], np.cumsum(tip)])
    major, minor, _ = _rank_tables(cum, size)
    out["tip_ones"] = np.int64(cum[-1]).tobytes(); out["tip_major"] = major.tobytes(); out["tip_minor"] = minor.tobytes()
    return out


def digest(sections):
    """{name: sha1 hex[:16]} -- what tests/golden/sdbg_golden.json stores"""
    return {k: hashlib.sha1(v).hexdigest()[:16] for k, v in sorted(sections.items())}
