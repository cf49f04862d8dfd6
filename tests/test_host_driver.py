"""The C++ host driver `megagta_b200 buildgraph` (drop-in for `megagta buildgraph`): option handling on the CPU,
and on the GPU the files it writes, read back with the reference-format reader and checked against the goldens
produced by the unmodified reference binary."""
import hashlib
import os
import subprocess

import pytest

from megagta_b200 import sdbg_io
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "megagta_b200", "bin", "megagta_b200")


def run(args, **kw):
    return subprocess.run([BIN, "buildgraph"] + args, capture_output=True, text=True, **kw)


def test_driver_is_built():
    assert os.path.exists(BIN), "run `python -c 'import __graft_entry__ as g; g.build()'`"


@pytest.mark.parametrize("args,msg", [
    (["--bogus", "1"], "uknown option"),                                         # options_description.cpp:69-70 (sic)
    (["-k", "31", "--host_mem", "1e9"], "No input file!"),                       # build_graph.cpp:53-55
    (["-k", "31", "--read_lib_file", "x"], "Please specify the host memory!"),   # :65-67
    (["--read_lib_file", "x", "--host_mem", "1e9", "--num_cpu_threads", "1"], "Number of CPU threads should be at least 2!"),
    (["--read_lib_file", "x", "--host_mem", "1e9", "--num_cpu_threads", "4", "--num_output_threads", "4"],
     "Number of output threads must be less than number of CPU threads!"),
])
def test_driver_rejects_bad_options_like_the_reference(args, msg):
    r = run(args)
    assert r.returncode == 1
    assert msg in r.stderr and "Usage: sdbg_builder read2sdbg" in r.stderr


def test_driver_reports_missing_library(tmp_path):
    r = run(["--read_lib_file", str(tmp_path / "nope"), "--host_mem", "1e9", "--num_cpu_threads", "2"])
    assert r.returncode == 1 and "[ERROR]" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["smoke_k31_m2", "smoke_k61_m2", "adversarial_k27_m3", "xander_k29_m1", "meta200k_k31_m2"])
def test_driver_writes_the_reference_files(case, golden, read_lib, tmp_path):
    g = golden["cases"][case]
    prefix, _ = read_lib(g["dataset"])
    out = str(tmp_path / "g")
    r = run(["-k", str(g["k"]), "-m", str(g["m"]), "--host_mem", "4e9", "--num_cpu_threads", "4", "--num_output_threads", "1",
             "--read_lib_file", prefix, "--output_prefix", out])
    assert r.returncode == 0, r.stderr
    hdr, stream, meta = sdbg_io.canonical(out)
    assert hdr["k"] == g["k"] and hdr["total_size"] == g["total_size"] and hdr["num_tips"] == g["num_tips"]
    assert hdr["large_multi"] == g["large_multi"]
    assert len(stream) == g["stream_bytes"]
    assert O.stream_hash(stream) == g["stream_hash"]
    assert O.meta_hash(meta) == g["meta_hash"]
    if g["m"] > 1:
        txt = open(out + ".counting").read()
        assert hashlib.sha256(txt.encode()).hexdigest()[:16] == g["counting_sha"]


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["smoke_k31_m2_assist", "smoke_k31_m1_assist", "adversarial_k27_m3_assist", "meta200k_k61_m2_assist"])
def test_driver_with_assist_seq_writes_the_reference_files(case, golden, read_lib, data_dir, tmp_path):
    """--assist_seq: the driver loads the FASTA like the reference (reversed, ACGTNacgtn map, .info),
    the kernels count the assist occurrences and treat their edges as solid; golden = the reference with --assist_seq."""
    import datasets
    g = golden["cases"][case]
    prefix, _ = read_lib(g["dataset"])
    fa = datasets.assist_fasta(g["dataset"], data_dir)
    assert datasets.md5(fa) == golden["datasets"][g["dataset"] + ".assist.fa"]
    out = str(tmp_path / "g")
    r = run(["-k", str(g["k"]), "-m", str(g["m"]), "--host_mem", "4e9", "--num_cpu_threads", "4", "--num_output_threads", "1",
             "--read_lib_file", prefix, "--assist_seq", fa, "--output_prefix", out])
    assert r.returncode == 0, r.stderr
    hdr, stream, meta = sdbg_io.canonical(out)
    assert hdr["total_size"] == g["total_size"] and hdr["num_tips"] == g["num_tips"] and hdr["large_multi"] == g["large_multi"]
    assert len(stream) == g["stream_bytes"]
    assert O.stream_hash(stream) == g["stream_hash"]
    assert O.meta_hash(meta) == g["meta_hash"]
    if g["m"] > 1:
        txt = open(out + ".counting").read()
        assert hashlib.sha256(txt.encode()).hexdigest()[:16] == g["counting_sha"]


def test_driver_reports_missing_assist_info(tmp_path):
    fa = tmp_path / "a.fa"
    fa.write_text(">x\nACGT\n")
    r = run(["--read_lib_file", str(tmp_path / "nope"), "--assist_seq", str(fa), "--host_mem", "1e9", "--num_cpu_threads", "2"])
    assert r.returncode == 1 and "[ERROR]" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["smoke_k31_m2_mercy", "adversarial_k27_m3_mercy", "xander_k29_m2_mercy"])
def test_driver_with_need_mercy_writes_the_reference_files(case, golden, read_lib, tmp_path):
    g = golden["cases"][case]
    prefix, _ = read_lib(g["dataset"])
    out = str(tmp_path / "g")
    r = run(["-k", str(g["k"]), "-m", str(g["m"]), "--host_mem", "4e9", "--num_cpu_threads", "4", "--num_output_threads", "1",
             "--read_lib_file", prefix, "--output_prefix", out, "--need_mercy"])
    assert r.returncode == 0, r.stderr
    hdr, stream, meta = sdbg_io.canonical(out)
    assert hdr["total_size"] == g["total_size"] and hdr["num_tips"] == g["num_tips"]
    assert O.stream_hash(stream) == g["stream_hash"] and O.meta_hash(meta) == g["meta_hash"]
    assert "Number mercy: %d" % g["num_mercy"] in r.stderr
