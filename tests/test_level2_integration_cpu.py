"""CPU: INTEGRATION.md "Level 2" as a program (integration/level2_build_graph.cpp): the reference's own option parser,
read loader and SdbgWriter, compiled from the UNMODIFIED reference sources around the C ABI (oracle/Makefile builds
oracle/_ref/megagta_level2 when /root/reference is present).  Checked without a GPU: it parses options and loads the read
library with the reference's code, then stops at mgta_ctx_create with the library's "no CUDA device" message (no fallback);
its stage-2 sink (integration/level2_sink.h), fed the oracle's record stream in deliveries (`megagta_level2 replay`), makes the
reference's SdbgWriter write files that read back as the same graph -- through the public write() and through the append_raw of
integration/sdbg_writer_append_raw.patch (megagta_level2_raw).  The run on a GPU is the opt-in test tests/test_gpu_level2.py."""
import os
import subprocess

import numpy as np
import pytest

import oracle_memo as OM
from megagta_b200 import sdbg_io
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LEVEL2 = os.path.join(ROOT, "oracle", "_ref", "megagta_level2")
LEVEL2_RAW = LEVEL2 + "_raw"       # the same program over a tree that carries integration/sdbg_writer_append_raw.patch


def need(binary, word=None):
    if not os.path.exists(binary) or (word and word not in open(binary, "rb").read().decode("latin1")):
        pytest.skip("%s not built (needs /root/reference at build time)" % os.path.relpath(binary, ROOT))


@pytest.mark.parametrize("args,msg", [(["--bogus", "1"], "uknown option"), (["-k", "31", "--host_mem", "1e9"], "No input file!"),
                                      (["-k", "31", "--read_lib_file", "x"], "Please specify the host memory!")])
@pytest.mark.parametrize("binary", [LEVEL2, LEVEL2_RAW])
def test_level2_keeps_the_reference_option_handling(binary, args, msg):
    need(binary)
    r = subprocess.run([binary, "buildgraph"] + args, capture_output=True, text=True, timeout=60)
    assert r.returncode == 1 and msg in r.stderr and "Usage: sdbg_builder read2sdbg" in r.stderr


@pytest.mark.parametrize("binary", [LEVEL2, LEVEL2_RAW])
def test_level2_loads_with_the_reference_loader_and_stops_at_the_device(binary, read_lib, tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    need(binary)
    prefix, rd = read_lib("smoke")
    r = subprocess.run([binary, "buildgraph", "-k", "31", "-m", "2", "--host_mem", "4e9", "--num_cpu_threads", "4", "--read_lib_file", prefix,
                        "--output_prefix", str(tmp_path / "g")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 1
    assert "%d reads, %d max read length, %d total bases" % (rd["n_reads"], rd["max_len"], int(rd["start"][-1])) in r.stderr
    assert "[ERROR]" in r.stderr and "no CUDA device (there is no CPU fallback)" in r.stderr
    assert not os.path.exists(str(tmp_path / "g") + ".sdbg_info")


@pytest.mark.parametrize("binary", [LEVEL2, LEVEL2_RAW])
@pytest.mark.parametrize("ds,k,m,per", [("smoke", 31, 2, 65536), ("smoke", 61, 2, 4096), ("adversarial", 21, 1, 1000), ("xander", 44, 2, 1)])
def test_level2_sink_hands_the_records_to_the_reference_writer(binary, read_lib, tmp_path, ds, k, m, per):
    """`megagta_level2 replay`: the oracle's record stream in deliveries of `per` buckets through the program's sink -- record
    by record through the public SdbgWriter::write, or (megagta_level2_raw) one append_raw per delivery -- and the reference's
    writer leaves files that read back as the same graph; the unmodified reference's LoadFromMultiFile loads them"""
    need(binary)
    _, rd = read_lib(ds)
    solid = OM.stage1(rd, k, m)[0] if m > 1 else None
    stream, meta, _ = OM.stage2(rd, k, m, solid)
    sf, mf, out = str(tmp_path / "stream"), str(tmp_path / "meta"), str(tmp_path / "g")
    open(sf, "wb").write(stream)
    np.ascontiguousarray(meta, dtype="<i8").tofile(mf)
    r = subprocess.run([binary, "replay", sf, mf, str(k), out, str(per)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-1000:]
    hdr, got_stream, got_meta = sdbg_io.canonical(out)
    assert hdr["k"] == k and hdr["num_threads"] == 1 and hdr["total_size"] == int(meta[:, 0].sum())
    assert hdr["num_tips"] == int(meta[:, 1].sum()) and hdr["large_multi"] == int(meta[:, 2].sum())
    assert got_stream == stream and np.array_equal(got_meta, meta)
    if O.have_ref() and "sdbgdump" in open(O.REF_BIN, "rb").read().decode("latin1"):
        from oracle import sdbg_oracle as SO
        ref = SO.ref_dump(O.REF_BIN, out, 1, str(tmp_path / "dump"))
        ref.pop("_load_seconds", None)
        want = SO.build(stream, np.asarray(meta), k, True)
        assert not [sec for sec in ref if want.get(sec) != ref[sec]]
    # a delivery whose table does not add up is refused
    bad = np.ascontiguousarray(meta, dtype="<i8").copy()
    bad[int(np.nonzero(meta[:, 0])[0][-1]), 0] += 1
    bad.tofile(mf)
    r = subprocess.run([binary, "replay", sf, mf, str(k), str(tmp_path / "bad"), str(per)], capture_output=True, text=True, timeout=300)
    assert r.returncode != 0


def test_append_raw_patch_applies_to_the_reference_writer(tmp_path):
    import shutil
    src = "/root/reference/src/sdbg_multi_io.h"
    if not os.path.exists(src) or not shutil.which("patch"):
        pytest.skip("the reference tree (or patch) is not present")
    os.makedirs(str(tmp_path / "src"))
    shutil.copy(src, str(tmp_path / "src" / "sdbg_multi_io.h"))
    r = subprocess.run(["patch", "-p1", "-i", os.path.join(ROOT, "integration", "sdbg_writer_append_raw.patch")], cwd=str(tmp_path),
                       capture_output=True, text=True)
    assert r.returncode == 0 and "FAILED" not in r.stdout and "fuzz" not in r.stdout, r.stdout + r.stderr
    assert "void append_raw(int tid" in open(str(tmp_path / "src" / "sdbg_multi_io.h")).read()


def test_level1_patch_applies_to_the_reference_driver(tmp_path):
    """integration/megagta_py_level1.patch (INTEGRATION.md Level 1: `--b200-bin` / $MEGAGTA_B200 swaps the executable of the
    buildlib, buildgraph and findstart steps) applies to the reference's megagta.py as it is, and the result still parses"""
    import shutil
    src = "/root/reference/src/megagta.py"
    if not os.path.exists(src) or not shutil.which("patch"):
        pytest.skip("the reference tree (or patch) is not present")
    os.makedirs(str(tmp_path / "src"))
    shutil.copy(src, str(tmp_path / "src" / "megagta.py"))
    r = subprocess.run(["patch", "-p1", "-i", os.path.join(ROOT, "integration", "megagta_py_level1.patch")], cwd=str(tmp_path),
                       capture_output=True, text=True)
    assert r.returncode == 0 and "FAILED" not in r.stdout and "fuzz" not in r.stdout, r.stdout + r.stderr
    text = open(str(tmp_path / "src" / "megagta.py")).read()
    compile(text, "megagta.py", "exec")
    assert text.count('sub_program("') == 3 and '[opt.bin_dir + "megagta", "buildgraph"]' not in text
