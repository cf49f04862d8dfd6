"""CPU: the Python restatements of `buildlib` and `findstart` (oracle/seqtools_oracle.py) against the UNMODIFIED reference
binary (buildlib: byte-identical files on the inputs of tests/test_gpu_buildlib.py) and against the committed digests of the
reference's findstart output (tests/golden/findstart_golden.json).  The GPU tests compare the product with the same
references; these pin the checker on the CPU."""
import importlib.util
import json
import os

import pytest

import datasets
import fastx_cases
from oracle import oracle as O
from oracle import seqtools_oracle as ST

HERE = os.path.dirname(os.path.abspath(__file__))


def _load(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(HERE, name + ".py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_buildlib_oracle_equals_the_reference_binary(tmp_path):
    if not O.have_ref():
        pytest.skip("oracle/_ref/megagta_ref not built")
    lib = _load("test_gpu_buildlib").write_inputs(str(tmp_path))
    ref = str(tmp_path / "ref")
    O.run_ref_buildlib(lib, ref)
    bin_bytes, info = ST.buildlib(lib)
    assert info == open(ref + ".lib_info").read()
    assert bin_bytes == open(ref + ".bin", "rb").read()


@pytest.mark.parametrize("name", sorted(fastx_cases.EDGE))
def test_buildlib_oracle_equals_the_reference_on_edge_case_files(tmp_path, name):
    """empty file, no final newline, header-only records, blank lines, CRLF FASTQ, multi-line FASTQ with '@' / '>' quality
    lines, truncated / missing quality, junk before the first header, FASTA and FASTQ records mixed"""
    if not O.have_ref():
        pytest.skip("oracle/_ref/megagta_ref not built")
    path = str(tmp_path / name)
    open(path, "wb").write(fastx_cases.EDGE[name])
    lib = str(tmp_path / "x.lib")
    open(lib, "w").write("edge case\nse %s\n" % path)
    ref = str(tmp_path / "ref")
    O.run_ref_buildlib(lib, ref)
    bin_bytes, info = ST.buildlib(lib)
    assert info == open(ref + ".lib_info").read()
    assert bin_bytes == open(ref + ".bin", "rb").read()


def test_pack_records_reproduces_the_in_tree_fixture():
    import numpy as np
    raw = open(os.path.join(datasets.GOLDEN_DIR, "xander.bin"), "rb").read()      # the reference's buildlib on its own test_reads.fa
    reads = ST.read_bin(os.path.join(datasets.GOLDEN_DIR, "xander.bin"))
    assert ST.pack_records([r.encode() for r in reads]) == raw
    assert np.frombuffer(ST.pack_records([b"", b"acgtn", b"N" * 33]), "<u4").tolist() == [0, 5, 0x1B800000, 33, 0xAAAAAAAA, 0xAAAAAAAA, 0x80000000]


@pytest.mark.parametrize("k_size,with_contigs", [(45, False), (30, True), (72, True), (44, False)])
def test_findstart_oracle_equals_the_reference_digests(tmp_path, k_size, with_contigs):
    T = _load("test_gpu_findstart")
    ref, binf, contigs = T.make_inputs(str(tmp_path))
    golden = json.load(open(os.path.join(datasets.GOLDEN_DIR, "findstart_golden.json")))
    lines = ST.find_seeds(ref, binf, k_size, contigs if with_contigs else None)
    assert T.digest(lines) == golden["k%d_contigs%d" % (k_size, int(with_contigs))]


GENE_DIR = "/root/reference/share/RDPTools/Xander_assembler/gene_resource"
GENES = ["nifH", "rplB", "nirK", "nirS", "nosZ", "nosZ_a2", "norB_cNor", "norB_qNor", "amoA_AOA", "amoA_AOB"]


@pytest.mark.parametrize("gene", GENES)
def test_findstart_oracle_equals_the_reference_on_its_own_gene_alignments(tmp_path, gene):
    """the reference's own aligned gene families (only in this container: the tree is absent on the GPU box) against the
    reference's in-tree reads (tests/golden/xander.bin = its test_reads.fa): the UNMODIFIED `findstart`, run live, and the
    oracle give the same seed lines (the reference shuffles them: compared sorted)"""
    import subprocess
    faa = os.path.join(GENE_DIR, gene, "ref_aligned.faa")
    if not os.path.exists(faa) or not O.have_ref():
        pytest.skip("the reference tree / oracle/_ref is not present")
    binf = os.path.join(datasets.GOLDEN_DIR, "xander.bin")
    for k_size in (45, 30):
        r = subprocess.run([O.REF_BIN, "findstart", faa, binf, str(k_size), "2"], capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-1000:]
        want = sorted(l for l in r.stdout.splitlines() if l)
        assert ST.find_seeds(faa, binf, k_size) == want
