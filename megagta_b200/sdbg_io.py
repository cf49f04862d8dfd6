"""Reader for the SdBG multi-file edge stream that `buildgraph` writes.

Format (reference: sdbg_multi_io.h:83-112 writer, :154-198 sdbg_info, :241-296 reader):
  <p>.sdbg_info  text: k, words_per_tip_label, num_buckets, num_threads, total_size, num_tips,
                 large_multi, then one row per bucket: bucket file_id byte_offset num_items num_tips num_large_mul
  <p>.sdbg.<i>   u16 records (+u16 multiplicity when (rec>>8)==255, +words_per_tip_label u32 when bit 5)

`canonical()` folds the files into the schedule-independent view SURVEY.md section 8(b) defines as the
parity object: the bucket-ordered byte stream and the per-bucket count triplets.
"""
import hashlib

import numpy as np

NUM_BUCKETS = 65536


def read_info(prefix):
    with open(prefix + ".sdbg_info") as f:
        lines = f.read().split("\n")
    hdr = {}
    for i, key in enumerate(["k", "words_per_tip_label", "num_buckets", "num_threads", "total_size", "num_tips",
                             "large_multi"]):
        name, val = lines[i].split()
        assert name == key, (name, key)
        hdr[key] = int(val)
    rows = np.array([list(map(int, ln.split())) for ln in lines[7:7 + hdr["num_buckets"]]], dtype=np.int64)
    assert rows.shape == (hdr["num_buckets"], 6)
    return hdr, rows


def canonical(prefix):
    """-> (hdr, stream bytes, meta int64[num_buckets,3])"""
    hdr, rows = read_info(prefix)
    wpt = hdr["words_per_tip_label"]
    files = [open("%s.sdbg.%d" % (prefix, i), "rb").read() for i in range(hdr["num_threads"])]
    parts = []
    for b in range(hdr["num_buckets"]):
        _, fid, off, n, tips, lm = rows[b]
        if fid == -1:
            continue
        sz = n * 2 + tips * 4 * wpt + lm * 2
        parts.append(files[fid][off:off + sz])
    return hdr, b"".join(parts), rows[:, 3:6].copy()


def stream_hash(stream):
    return hashlib.sha256(stream).hexdigest()[:16]


def meta_hash(meta):
    m = np.asarray(meta)
    txt = "".join("%d %d %d %d\n" % (b, m[b, 0], m[b, 1], m[b, 2]) for b in range(m.shape[0]))
    return hashlib.sha256(txt.encode()).hexdigest()[:16]


def decode_stream(stream, words_per_tip_label):
    """Decode records -> list of (w, last, tip, multiplicity, tip_label tuple or None). For small streams."""
    out = []
    a = np.frombuffer(stream, dtype="<u2")
    i = 0
    while i < len(a):
        rec = int(a[i]); i += 1
        mult = rec >> 8
        if mult == 255:
            mult = int(a[i]); i += 1
        label = None
        if rec >> 5 & 1:
            label = tuple(int(a[i + 2 * j]) | int(a[i + 2 * j + 1]) << 16 for j in range(words_per_tip_label))
            i += 2 * words_per_tip_label
        out.append((rec & 15, rec >> 4 & 1, rec >> 5 & 1, mult, label))
    return out
