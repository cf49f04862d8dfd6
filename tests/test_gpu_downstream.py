"""Downstream acceptance: the UNMODIFIED reference `denovo` (assembler.cpp -> SuccinctDBG::
LoadFromMultiFile, succinct_dbg.cpp:595-723: opens every <p>.sdbg.<i>, walks the sdbg_info rows, builds rank/select) loads
the files `megagta_b200 buildgraph` wrote -- one file per GPU -- and assembles the same contigs as from the files the
reference's own buildgraph wrote.  oracle/_ref/megagta_ref is the reference compiled by oracle/Makefile (checker only)."""
import os
import subprocess

import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "megagta_b200", "bin", "megagta_b200")
COMP = str.maketrans("ACGT", "TGCA")


def contigs(path):
    """multiset of contig sequences, each in its canonical orientation (denovo's order and strand depend on threading)"""
    seqs, cur = [], []
    for line in open(path):
        if line.startswith(">"):
            if cur:
                seqs.append("".join(cur))
            cur = []
        else:
            cur.append(line.strip())
    if cur:
        seqs.append("".join(cur))
    return sorted(min(s, s.translate(COMP)[::-1]) for s in seqs)


def denovo(prefix, min_contig):
    # one thread: with -t 4 the reference's own bubble / tip removal is not deterministic (5 different contig sets in 8 runs
    # on its own meta200k files: equal-support alleles are resolved in thread-arrival order)
    r = subprocess.run([O.REF_BIN, "denovo", "-s", prefix, "-o", prefix, "-t", "1", "--min_standalone", "400", "--max_tip_len", "150",
                        "--min_contig", str(min_contig)], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    return contigs(prefix + ".contigs.fa"), r.stderr


@pytest.mark.parametrize("ds,k,m,gpus", [("smoke", 31, 2, 1), ("xander", 29, 1, 1), ("meta200k", 31, 2, 1), ("meta200k", 41, 2, 2),
                                          ("smoke", 31, 2, 2)])
def test_reference_denovo_accepts_our_files(read_lib, tmp_path, ds, k, m, gpus):
    import torch
    if not O.have_ref():
        pytest.skip("oracle/_ref/megagta_ref not built")
    if "denovo" not in subprocess.run([O.REF_BIN], capture_output=True, text=True).stderr:
        pytest.skip("oracle/_ref/megagta_ref predates the downstream build (rebuild with make -C oracle)")
    if torch.cuda.device_count() < gpus:
        pytest.skip("needs %d GPUs" % gpus)
    prefix, _ = read_lib(ds)
    ours, ref = str(tmp_path / "ours"), str(tmp_path / "ref")
    r = subprocess.run([BIN, "buildgraph", "-k", str(k), "-m", str(m), "--host_mem", "4e9", "--num_cpu_threads", "4", "--num_output_threads", "1",
                        "--read_lib_file", prefix, "--output_prefix", ours], capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, MGTA_NUM_GPUS=str(gpus)))
    assert r.returncode == 0, r.stderr[-3000:]
    O.run_ref_buildgraph(prefix, ref, k, m, threads=4)
    for g in range(gpus):
        assert os.path.exists("%s.sdbg.%d" % (ours, g))
    mine, log_mine = denovo(ours, k + 6)
    theirs, log_ref = denovo(ref, k + 6)
    assert len(theirs) > 0
    assert mine == theirs
    # the loader's own summary of the graph it built must agree too (succinct_dbg.cpp logs sizes; assembler.cpp:94 totals)
    tot = [ln.split("]", 1)[1].strip() for ln in log_ref.splitlines() if "Total length" in ln]
    tot_mine = [ln.split("]", 1)[1].strip() for ln in log_mine.splitlines() if "Total length" in ln]
    assert tot and tot == tot_mine
