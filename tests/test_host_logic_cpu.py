"""CPU: the HOST-side logic of the C++ drivers (megagta_b200/csrc/host/*), compiled as it is into a test harness
(tests/cpu/driver_host.cpp) and checked without a GPU:

  * `load_read_lib` / `append_assist` (reversed, bit-contiguous packing: read_lib_functions-inl.h:233-261,
    sequence_package.h:247-252,341-367, s1.cpp:104-134) against the oracle's numpy loader, bit for bit;
  * the record-file writer + sdbg_info (sdbg_multi_io.h:83-112,154-198): the oracle's record stream written through the
    driver's own `sink` into 1 and 3 files, read back by the reference-format reader and -- when oracle/_ref is built -- by
    the UNMODIFIED reference's SuccinctDBG::LoadFromMultiFile (`sdbgdump`);
  * the kseq-rule FASTA / FASTQ reader (kseq.h:168-207) and findstart's model k-mer rules (prot_kmer_generator.h:58-137)
    against oracle/seqtools_oracle.py, which is pinned on the reference binary (tests/test_seqtools_oracle.py)."""
import ctypes
import gzip
import importlib.util
import os
import subprocess

import numpy as np
import pytest

import datasets
import fastx_cases
import oracle_memo as OM
from megagta_b200 import sdbg_io
from oracle import oracle as O
from oracle import seqtools_oracle as ST

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "megagta_b200", "csrc", "host")
SRC = os.path.join(ROOT, "tests", "cpu", "driver_host.cpp")
OUT = os.path.join(ROOT, "tests", "cpu", "_build", "libdriver_host.so")
LIBDIR = os.path.join(ROOT, "megagta_b200", "lib")


@pytest.fixture(scope="session")
def host():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    deps = [SRC] + [os.path.join(HOST, f) for f in os.listdir(HOST)] + [os.path.join(ROOT, "include", "mgta_cuda.h")]
    if not os.path.exists(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in deps):
        subprocess.run(["g++", "-O2", "-std=c++17", "-fopenmp", "-pthread", "-shared", "-fPIC", "-I/usr/local/cuda/include", "-o", OUT, SRC,
                        "-L" + LIBDIR, "-lmgta_cuda", "-L/usr/local/cuda/lib64", "-lcudart", "-lnccl", "-lz",
                        "-Wl,-rpath," + LIBDIR, "-Wl,-rpath,/usr/local/cuda/lib64"], check=True)
    lib = ctypes.CDLL(OUT)
    for f in (lib.hd_load_reads, lib.hd_fastx, lib.hd_model_kmers):
        f.restype = ctypes.c_int64
    return lib


def _load(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "tests", name + ".py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def load_reads(host, prefix, assist=""):
    seq, start = ctypes.POINTER(ctypes.c_uint32)(), ctypes.POINTER(ctypes.c_uint64)()
    n_words, n_short, max_len = ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_int()
    n = host.hd_load_reads(prefix.encode(), assist.encode(), 4, ctypes.byref(seq), ctypes.byref(n_words), ctypes.byref(start),
                           ctypes.byref(max_len), ctypes.byref(n_short))
    s = np.ctypeslib.as_array(seq, (n_words.value,)).copy()
    t = np.ctypeslib.as_array(start, (n + 1,)).copy()
    host.hd_free(seq)
    host.hd_free(start)
    return dict(seq=s, start=t, n_reads=int(n), max_len=max_len.value), int(n_short.value)


@pytest.mark.parametrize("ds", ["tiny", "smoke", "adversarial", "xander", "meta200k"])
def test_driver_loader_packs_the_reads_like_the_oracle(host, read_lib, ds):
    prefix, rd = read_lib(ds)
    got, n_short = load_reads(host, prefix)
    assert got["n_reads"] == rd["n_reads"] == n_short and got["max_len"] == rd["max_len"]
    assert np.array_equal(got["start"], rd["start"])
    n = int(rd["start"][-1]) // 16 + 1
    assert len(got["seq"]) == n and np.array_equal(got["seq"], rd["seq"][:n])


@pytest.mark.parametrize("ds,assist", [("tiny", False), ("smoke", True), ("adversarial", True), ("xander", False)])
def test_loaders_equal_the_reference_sequence_package(host, read_lib, data_dir, tmp_path, ds, assist):
    """the read set as the UNMODIFIED reference holds it in memory after its own s1_read_input_prepare (`megagta_ref readsdump`:
    ReadBinaryLibs with is_reverse = true, --assist_seq appended; cx1_read2sdbg_s1.cpp:96-134) = what the oracle's numpy loader
    builds = what the driver's C++ loader hands to mgta_set_reads, bit for bit"""
    if not O.have_ref() or "readsdump" not in open(O.REF_BIN, "rb").read().decode("latin1"):
        pytest.skip("oracle/_ref/megagta_ref with readsdump not built")
    from oracle import sdbg_oracle as SO
    prefix, rd = read_lib(ds)
    fa = datasets.assist_fasta(ds, data_dir) if assist else ""
    dump = str(tmp_path / "reads.dump")
    r = subprocess.run([O.REF_BIN, "readsdump", prefix, fa, dump], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-1000:]
    d = SO.read_dump(dump)
    n_reads, n_short, max_len, bases = (int(x) for x in np.frombuffer(d["hdr"], np.int64))
    ref_seq, ref_start = np.frombuffer(d["packed_seq"], np.uint32), np.frombuffer(d["start_idx"], np.uint64)
    exp, n0 = O.with_assist(rd, fa) if assist else (rd, rd["n_reads"])
    got, got_short = load_reads(host, prefix, fa)
    n = bases // 16 + 1
    assert len(ref_seq) == n                                          # the n_words the ABI is given
    for name, x in (("oracle", exp), ("driver", got)):
        assert x["n_reads"] == n_reads and x["max_len"] == max_len, name
        assert np.array_equal(x["start"], ref_start), name
        assert np.array_equal(x["seq"][:n], ref_seq), name
    assert n0 == n_short == got_short


ASSIST_EDGE_FILES = {
    "multi_line_fasta": ">c0 some text\nACGTNacgtnACGTACGTACGTACGTACGTACGTACGT\nacgtacgtacgtnnnnACGT\nAC\n>c1\nGGGGCCCCAAAATTTTGGGGCCCCAAAATTTTG\n",
    "crlf_fasta": ">c0\r\nACGTACGTACGTACGTACGTACGTACGTACGTAC\r\nGGGTTT\r\n>c1\r\nTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTT\r\n",
    "fastq_with_marker_qualities": "@a\nACGTACGTACGTACGTACGTACGTAAAA\nCCCCGGGGTTTTACGTACGTACGTACGT\n+a\n@IIIIIIIIIIIIIIIIIIIIIIIIIII\n"
                                   ">IIIIIIIIIIIIIIIIIIIIIIIIIII\n@b\nACGTACGTACGTACGTACGTACGTACGTAA\n+\nIIIIIIIIIIIIIIIIIIIIIIIIIIIIII\n",
    "blank_lines": "\n\n>a\n\nACGTACGTACGTACGTACGTACGTACGTACGT\n\nACGT\n\n>b\nACGTACGTACGTACGTACGTACGTACGTACGTTT\n\n",
    "no_final_newline": ">a\nACGTACGTACGTACGTACGTACGTACGTACGT\n>b\nACGTACGTACGTACGTACGTACGTACGTACGTTT",
    "empty_records": ">a\n>b\nACGTACGTACGTACGTACGTACGTACGTACGTTT\n>c\n",
}


@pytest.mark.parametrize("name", sorted(ASSIST_EDGE_FILES))
def test_assist_parser_equals_the_reference_on_edge_case_files(host, read_lib, tmp_path, name):
    """--assist_seq files the reference reads through kseq (s1.cpp:120-134): multi-line and CRLF FASTA, FASTQ whose quality
    lines start with '@' / '>', blank lines, no final newline, empty records -- the driver's append_assist against the
    reference's own SequencePackage (`megagta_ref readsdump`)"""
    if not O.have_ref() or "readsdump" not in open(O.REF_BIN, "rb").read().decode("latin1"):
        pytest.skip("oracle/_ref/megagta_ref with readsdump not built")
    from oracle import sdbg_oracle as SO
    prefix, rd = read_lib("tiny")
    fa = str(tmp_path / (name + ".fx"))
    open(fa, "w", newline="").write(ASSIST_EDGE_FILES[name])
    open(fa + ".info", "w").write("0 0\n")                           # only its presence matters to either loader
    dump = str(tmp_path / "reads.dump")
    r = subprocess.run([O.REF_BIN, "readsdump", prefix, fa, dump], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-1000:]
    d = SO.read_dump(dump)
    ref_seq, ref_start = np.frombuffer(d["packed_seq"], np.uint32), np.frombuffer(d["start_idx"], np.uint64)
    got, n_short = load_reads(host, prefix, fa)
    assert n_short == rd["n_reads"] and got["n_reads"] == len(ref_start) - 1 > n_short
    assert np.array_equal(got["start"], ref_start)
    n = int(ref_start[-1]) // 16 + 1
    assert np.array_equal(got["seq"][:n], ref_seq[:n])


def test_driver_loader_on_ragged_and_gzipped_libraries(host, tmp_path):
    """lengths 0..90 in random order (every alignment of a read against the word grid, empty reads, reads inside one word),
    once as a plain .bin (mapped) and once gzip'ed (inflated through zlib, as the reference reads it)"""
    from megagta_b200 import synth
    rng = np.random.default_rng(17)
    reads = [rng.integers(0, 4, int(n), dtype=np.uint8) for n in rng.integers(0, 91, 6000)]
    reads += [rng.integers(0, 4, n, dtype=np.uint8) for n in (0, 0, 1, 15, 16, 17, 31, 32, 33, 1000, 4097)]
    prefix = str(tmp_path / "ragged")
    synth.write_variable_reads(prefix, reads)
    rd = O.load_read_lib(prefix)
    got, _ = load_reads(host, prefix)
    n = int(rd["start"][-1]) // 16 + 1
    assert got["n_reads"] == len(reads) and got["max_len"] == 4097
    assert np.array_equal(got["start"], rd["start"]) and np.array_equal(got["seq"], rd["seq"][:n])
    gz = str(tmp_path / "ragged_gz")
    with gzip.open(gz + ".bin", "wb") as f:
        f.write(open(prefix + ".bin", "rb").read())
    open(gz + ".lib_info", "w").write(open(prefix + ".lib_info").read())
    got_gz, _ = load_reads(host, gz)
    assert np.array_equal(got_gz["start"], got["start"]) and np.array_equal(got_gz["seq"], got["seq"])


@pytest.mark.parametrize("cut,msg", [(-3, "truncated record in"), (-8, "truncated record in"), (2, "truncated record header"), (None, ".lib_info says")])
def test_driver_loader_rejects_damaged_libraries(host, read_lib, tmp_path, cut, msg):
    prefix, _ = read_lib("tiny")
    bad = str(tmp_path / "bad")
    raw = open(prefix + ".bin", "rb").read()
    open(bad + ".bin", "wb").write(raw if cut is None else raw[:cut] if cut < 0 else raw + b"\x07" * cut)
    info = open(prefix + ".lib_info").read()
    if cut is None:
        first, rest = info.split("\n", 1)
        info = "%d %d\n%s" % (int(first.split()[0]) + 1, int(first.split()[1]), rest)
    open(bad + ".lib_info", "w").write(info)
    code = ("import ctypes; l = ctypes.CDLL(%r); p = ctypes.c_void_p(); q = ctypes.c_void_p(); a = ctypes.c_uint64(); b = ctypes.c_uint64(); "
            "m = ctypes.c_int(); l.hd_load_reads(%r, b'', 2, ctypes.byref(p), ctypes.byref(a), ctypes.byref(q), ctypes.byref(m), ctypes.byref(b))"
            % (OUT, bad.encode()))
    r = subprocess.run([os.sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert r.returncode == 1 and "[ERROR]" in r.stderr and msg in r.stderr, r.stderr


@pytest.mark.parametrize("ds", ["smoke", "adversarial"])
def test_driver_appends_assist_reads_like_the_oracle(host, read_lib, data_dir, ds):
    prefix, rd = read_lib(ds)
    fa = datasets.assist_fasta(ds, data_dir)
    exp, n0 = O.with_assist(rd, fa)
    got, n_short = load_reads(host, prefix, fa)
    assert n_short == n0 and got["n_reads"] == exp["n_reads"] and got["max_len"] == exp["max_len"]
    assert np.array_equal(got["start"], exp["start"])
    n = int(exp["start"][-1]) // 16 + 1
    assert np.array_equal(got["seq"][:n], exp["seq"][:n])


def write_graph(host, prefix, k, stream, meta, cuts, per_delivery, true_meta=None):
    """the oracle's records through the driver's sink: deliveries of `per_delivery` buckets, files cut at `cuts`"""
    m = np.ascontiguousarray(meta, dtype=np.int64)
    wpt = (2 * k + 31) // 32
    t = m if true_meta is None else np.asarray(true_meta, dtype=np.int64)      # the bytes really delivered
    size = t[:, 0] * 2 + t[:, 2] * 2 + t[:, 1] * 4 * wpt
    bounds = sorted(set(list(range(0, 65536, per_delivery)) + list(cuts) + [65536]))
    b0 = np.array(bounds[:-1], dtype=np.int32)
    b1 = np.array(bounds[1:], dtype=np.int32)
    csum = np.concatenate([[0], np.cumsum(size)])
    d_bytes = (csum[b1] - csum[b0]).astype(np.uint64)
    cut = np.array(list(cuts) + [65536], dtype=np.int32)
    buf = np.frombuffer(stream, dtype=np.uint8)
    return host.hd_write_graph(prefix.encode(), k, len(cuts), cut.ctypes.data_as(ctypes.c_void_p), len(b0), b0.ctypes.data_as(ctypes.c_void_p),
                               b1.ctypes.data_as(ctypes.c_void_p), buf.ctypes.data_as(ctypes.c_void_p),
                               d_bytes.ctypes.data_as(ctypes.c_void_p), m.ctypes.data_as(ctypes.c_void_p))


@pytest.mark.parametrize("ds,k,m,cuts,per", [("smoke", 31, 2, [0], 65536), ("smoke", 31, 2, [0, 20000, 47001], 4096),
                                            ("adversarial", 21, 1, [0, 30000], 1000), ("xander", 44, 2, [0, 1, 65535], 777)])
def test_driver_writer_files_read_back_as_the_same_graph(host, read_lib, tmp_path, ds, k, m, cuts, per):
    _, rd = read_lib(ds)
    solid = OM.stage1(rd, k, m)[0] if m > 1 else None
    stream, meta, totals = OM.stage2(rd, k, m, solid)
    out = str(tmp_path / "g")
    total = write_graph(host, out, k, stream, meta, cuts, per)
    assert total == int(meta[:, 0].sum())
    hdr, got_stream, got_meta = sdbg_io.canonical(out)                 # the reference-format reader of the repo
    assert hdr["k"] == k and hdr["num_threads"] == len(cuts) and hdr["total_size"] == total
    assert hdr["num_tips"] == int(meta[:, 1].sum()) and hdr["large_multi"] == int(meta[:, 2].sum())
    assert got_stream == stream and np.array_equal(got_meta, meta)
    for f in range(len(cuts)):
        assert os.path.exists("%s.sdbg.%d" % (out, f))               # the reader opens every file, even an empty one
    # rows of empty buckets carry file_id -1 (sdbg_multi_io.h:356-358)
    rows = [l.split() for l in open(out + ".sdbg_info").read().splitlines()[7:]]
    assert len(rows) == 65536 and all((r[1] == "-1") == (r[3] == "0") for r in rows)
    if O.have_ref() and "sdbgdump" in open(O.REF_BIN, "rb").read().decode("latin1"):
        # the UNMODIFIED reference loads the files (SuccinctDBG::LoadFromMultiFile) and holds the graph the oracle describes
        from oracle import sdbg_oracle as SO
        ref = SO.ref_dump(O.REF_BIN, out, 1, str(tmp_path / "dump"))
        ref.pop("_load_seconds", None)
        want = SO.build(stream, np.asarray(meta), k, True)
        bad = [sec for sec in ref if want.get(sec) != ref[sec]]
        assert not bad, bad


def test_writer_refuses_a_table_that_does_not_add_up(host, read_lib, tmp_path):
    _, rd = read_lib("tiny")
    stream, meta, _ = OM.stage2(rd, 21, 1, None)
    bad = meta.copy()
    bad[int(np.nonzero(meta[:, 0])[0][0]), 0] += 1
    assert write_graph(host, str(tmp_path / "g"), 21, stream, bad, [0], 65536, true_meta=meta) == -2


def fastx(host, path):
    out, off = ctypes.c_char_p(), ctypes.POINTER(ctypes.c_uint64)()
    n = host.hd_fastx(path.encode(), ctypes.byref(out), ctypes.byref(off))
    o = np.ctypeslib.as_array(off, (n + 1,)).copy()
    raw = ctypes.string_at(out, int(o[-1]))
    host.hd_free(out)
    host.hd_free(off)
    return [raw[int(a):int(b)] for a, b in zip(o[:-1], o[1:])]


def test_fastx_reader_follows_the_kseq_rules(host, tmp_path):
    d = str(tmp_path)
    _load("test_gpu_buildlib").write_inputs(d)                         # multi-line, CRLF, empty lines, gzip, '@' in qualities
    for f in ("a.fa", "b.fq.gz", "p_1.fq", "p_2.fq", "i.fa"):
        path = os.path.join(d, f)
        want = ST.fastx_sequences(path)
        assert len(want) > 1000 and fastx(host, path) == want
    edge = fastx_cases.EDGE
    for name, data in edge.items():
        path = os.path.join(d, name)
        open(path, "wb").write(data)
        assert fastx(host, path) == ST.fastx_sequences(path), name
    path = os.path.join(d, "big.fa.gz")                                # records that straddle the reader's 1 MiB buffer
    rng = np.random.default_rng(3)
    with gzip.open(path, "wb") as f:
        for i in range(40):
            s = bytes(np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, int(rng.integers(1, 200000)))])
            f.write(b">r%d\n" % i + b"\n".join(s[o:o + 70] for o in range(0, len(s), 70)) + b"\n")
    assert fastx(host, path) == ST.fastx_sequences(path)


@pytest.mark.parametrize("k", [10, 12, 13, 15, 24])
def test_model_kmers_follow_the_generator_rules(host, tmp_path, k):
    ref, _, _ = _load("test_gpu_findstart").make_inputs(str(tmp_path))
    rows, text = ctypes.POINTER(ctypes.c_uint64)(), ctypes.c_char_p()
    n = host.hd_model_kmers(ref.encode(), k, ctypes.byref(rows), ctypes.byref(text))
    r = np.ctypeslib.as_array(rows, (max(n, 1), 3))[:n].copy()
    t = ctypes.string_at(text, n * k).decode()
    host.hd_free(rows)
    host.hd_free(text)
    got = {}
    for i in range(n):                                                 # the first occurrence of a k-mer keeps its position
        got.setdefault(t[i * k:(i + 1) * k].upper(), int(r[i, 2]))
    want = ST.model_kmers(ref, k)
    assert n > 500 and got == want
    # the packed key: 5 bits per residue, residues 0..11 in word 0 (first residue most significant), 12.. in word 1
    code = {c: i for i, c in enumerate(ST.RESIDUES)}
    for i in range(0, n, 97):
        km = t[i * k:(i + 1) * k].upper()
        w0 = w1 = 0
        for j, c in enumerate(km):
            if j < 12:
                w0 = (w0 << 5) | code[c]
            else:
                w1 = (w1 << 5) | code[c]
        assert (int(r[i, 0]), int(r[i, 1])) == (w0, w1)


@pytest.mark.parametrize("gene", ["nifH", "rplB", "nirK", "nosZ"])
def test_model_kmers_on_the_reference_gene_alignments(host, gene):
    """findstart's host rules on the reference's own aligned gene families (present in this container only)"""
    faa = "/root/reference/share/RDPTools/Xander_assembler/gene_resource/%s/ref_aligned.faa" % gene
    if not os.path.exists(faa):
        pytest.skip("the reference tree is not present")
    for k in (10, 15):
        rows, text = ctypes.POINTER(ctypes.c_uint64)(), ctypes.c_char_p()
        n = host.hd_model_kmers(faa.encode(), k, ctypes.byref(rows), ctypes.byref(text))
        r = np.ctypeslib.as_array(rows, (max(n, 1), 3))[:n].copy()
        t = ctypes.string_at(text, n * k).decode()
        host.hd_free(rows)
        host.hd_free(text)
        got = {}
        for i in range(n):
            got.setdefault(t[i * k:(i + 1) * k].upper(), int(r[i, 2]))
        assert n > 10000 and got == ST.model_kmers(faa, k)
