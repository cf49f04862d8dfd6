"""oracle/seqtools_oracle.py -- TEST INFRASTRUCTURE ONLY (checker), never imported by the product package.

Plain-Python restatements of the two streaming sub-programs next to the graph build, pinned on the
unmodified reference binary (tests/test_seqtools_oracle.py, tests/golden/findstart_golden.json):

  buildlib   ReadAndWriteMultipleLibs (read_lib_functions-inl.h:119-226) over kseq's record rules (kseq.h:168-207), the
             character map of SequencePackage (sequence_package.h:67-69) and the record layout of
             SequenceManager::WriteBinarySequences (sequence_manager.cpp:375-410)
  findstart  find_start / ProcessSequenceMulti (fast_kmer_filter.cpp:49-218): protein k-mers of the aligned reference with the
             model-only rules of ProtKmerGenerator (prot_kmer_generator.h:58-137), reads translated in three frames on both
             strands (sequence/Codon.C:8-90), one seed line per distinct nucleotide k-mer
"""
import gzip

import numpy as np

DNA = {c: v for c, v in zip("ACGTNacgtn", [0, 1, 2, 3, 2, 0, 1, 2, 3, 2])}
RESIDUES = "ARNDCQEGHILKMFPSTWYV"
CODON_AA = "KNKNTTTTRSRSIIMIQHQHPPPPRRRRLLLLEDEDAAAAGGGGVVVV*Y*YSSSS*CWCLFLF"     # index 16 b0 + 4 b1 + b2, ACGT = 0123


# ----------------------------------------------------------------------------------------------- kseq record rules
def fastx_sequences(path):
    """sequences of a FASTA / FASTQ file (gzip or plain) as kseq_read returns them (kseq.h:168-207)"""
    raw = open(path, "rb").read()
    data = gzip.decompress(raw) if raw[:2] == b"\x1f\x8b" else raw
    pos, n, last = 0, len(data), 0

    def rest_of_line(acc):                     # ks_getuntil2(KS_SEP_LINE, append): up to '\n', a trailing '\r' of the string dropped
        nonlocal pos
        e = data.find(b"\n", pos)
        if e < 0:
            e = n
        acc += data[pos:e]
        pos = min(n, e + 1)
        if len(acc) > 1 and acc[-1:] == b"\r":
            del acc[-1]

    out = []
    while True:
        if last == 0:                          # jump to the next header character
            while pos < n and data[pos:pos + 1] not in (b">", b"@"):
                pos += 1
            if pos >= n:
                break
            last = data[pos]
            pos += 1
        start = pos                            # name up to white space; the rest of the line is the comment
        while pos < n and not data[pos:pos + 1].isspace():
            pos += 1
        if pos >= n and pos == start:
            break
        c = data[pos:pos + 1]
        pos += 1
        if c != b"\n" and pos <= n:
            e = data.find(b"\n", pos)
            pos = n if e < 0 else e + 1
        seq = bytearray()
        c = -1
        while pos < n:
            c = data[pos]
            pos += 1
            if c in (0x3E, 0x2B, 0x40):        # '>', '+', '@'
                break
            if c == 0x0A:
                c = -1
                continue
            seq.append(c)
            rest_of_line(seq)
            c = -1
        if c in (0x3E, 0x40):
            last = c
        if c != 0x2B:                          # FASTA record
            out.append(bytes(seq))
            if c == -1:
                break                          # end of file
            continue
        e = data.find(b"\n", pos)              # rest of the '+' line
        if e < 0:
            break                              # no quality string: kseq_read() < 0 ends the file for the reference
        pos = e + 1
        qual = bytearray()
        while True:
            if pos >= n:
                break
            rest_of_line(qual)
            if len(qual) >= len(seq):
                break
        last = 0
        if len(qual) != len(seq):
            break
        out.append(bytes(seq))
    return out


# ----------------------------------------------------------------------------------------------- buildlib
def pack_records(seqs):
    """<X>.bin records: per read u32 length + ceil(len / 16) words, first base in bits 31..30 (sequence_manager.cpp:375-410)"""
    parts = []
    for s in seqs:
        n = len(s)
        codes = np.array([DNA.get(chr(c), 0) for c in s] + [0] * ((-n) % 16), dtype=np.uint32).reshape(-1, 16)
        words = (codes << (30 - 2 * np.arange(16, dtype=np.uint32))).sum(axis=1, dtype=np.uint32) if n else np.zeros(0, np.uint32)
        parts.append(np.concatenate([np.array([n], np.uint32), words.astype(np.uint32)]))
    return np.concatenate(parts).astype("<u4").tobytes() if parts else b""


def buildlib(lib_file):
    """-> (bytes of <P>.bin, text of <P>.lib_info) for a read-library list (read_lib_functions-inl.h:119-226)"""
    lines = open(lib_file).read().split("\n")
    bin_parts, libs, total_reads, total_bases, i = [], [], 0, 0, 0
    while i + 1 < len(lines) and (lines[i] != "" or i + 1 < len(lines) - 1):
        metadata, fields = lines[i], lines[i + 1].split()
        i += 2
        if not fields:
            break
        typ = fields[0]
        if typ == "pe":
            a, b = fastx_sequences(fields[1]), fastx_sequences(fields[2])
            assert len(a) == len(b)
            seqs = [s for pair in zip(a, b) for s in pair]
        else:
            seqs = fastx_sequences(fields[1])
        bin_parts.append(pack_records(seqs))
        start = total_reads
        total_reads += len(seqs)
        total_bases += sum(len(s) for s in seqs)
        libs.append((metadata, start, total_reads - 1, max([len(s) for s in seqs] + [0]), typ != "se"))
    info = "%d %d\n" % (total_bases, total_reads)
    for metadata, a, b, mx, pe in libs:
        info += "%s\n%d %d %d %s\n" % (metadata, a, b, mx, "pe" if pe else "se")
    return b"".join(bin_parts), info


# ----------------------------------------------------------------------------------------------- findstart
def model_kmers(ref_faa, k):
    """{protein k-mer (upper case residues) -> model position of its first occurrence} (prot_kmer_generator.h:58-137 with
    model_only = true; HashSetST::insert keeps the first)"""
    out = {}
    for seq in fastx_sequences(ref_faa):
        win, position = [], 1
        for c in seq.decode("latin1"):
            if c.islower() or c in "-Xx":
                if c in "-X":
                    position += 1
                win = []
                continue
            if c == "." or c == "*" or c.upper() not in RESIDUES:
                continue
            win.append(c.upper())
            position += 1
            if len(win) >= k:
                out.setdefault("".join(win[-k:]), position - k)
    return out


def read_bin(path):
    raw = np.frombuffer(open(path, "rb").read(), dtype="<u4")
    reads, p = [], 0
    while p < len(raw):
        n = int(raw[p])
        w = raw[p + 1:p + 1 + (n + 15) // 16]
        codes = ((w[:, None] >> (30 - 2 * np.arange(16, dtype=np.uint32))) & 3).reshape(-1)[:n]
        reads.append("".join("ACGT"[c] for c in codes))
        p += 1 + (n + 15) // 16
    return reads


def find_seeds(ref_faa, bin_path, k_size, contigs=None):
    """sorted seed lines of `megagta findstart` (fast_kmer_filter.cpp:49-218; the reference shuffles them)"""
    k = k_size // 3
    model = model_kmers(ref_faa, k)
    comp = str.maketrans("ACGT", "TGCA")
    seqs = read_bin(bin_path)
    if contigs:
        seqs += ["".join("ACGT"[DNA.get(chr(c), 0)] for c in s) for s in fastx_sequences(contigs)]
    seeds = {}
    idx = {c: i for i, c in enumerate("ACGT")}
    for s in seqs:
        if len(s) < k_size:
            continue
        for strand in (s, s.translate(comp)[::-1]):
            for frame in range(3):
                aa = "".join(CODON_AA[16 * idx[strand[p]] + 4 * idx[strand[p + 1]] + idx[strand[p + 2]]]
                             for p in range(frame, len(strand) - 2, 3))
                for j in range(len(aa) - k + 1):
                    pos = model.get(aa[j:j + k])
                    if pos is not None:
                        at = 3 * j + frame
                        seeds.setdefault(strand[at:at + k_size], (aa[j:j + k].lower(), pos))
    return sorted("dump_gene_name\tdump_seq_name\tdump\t%s\ttrue\t1\t%s\t%d" % (n, p, m) for n, (p, m) in seeds.items())
