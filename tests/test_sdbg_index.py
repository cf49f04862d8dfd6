"""SdBG load + rank/select build.

CPU: the numpy restatement (oracle/sdbg_oracle.py) against digests of the arrays the UNMODIFIED reference holds after
SuccinctDBG::LoadFromMultiFile (tests/golden/sdbg_golden.json, tests/golden/make_sdbg_golden.py).
GPU: the device build (mgta_sdbg_*, megagta_b200/csrc/sdbg_index.cu) against the restatement, the golden digests and --
when oracle/_ref/megagta_ref is present -- a live dump of the reference's members after it loaded OUR files."""
import json
import os
import subprocess

import numpy as np
import pytest

import datasets
from oracle import oracle as O
from oracle import sdbg_oracle as SO

CASES = [("tiny", 25, 2), ("smoke", 31, 2), ("smoke", 21, 1), ("smoke", 61, 2), ("adversarial", 31, 2), ("adversarial", 21, 1),
         ("xander", 29, 1), ("meta200k", 31, 2)]
CPU_CASES = CASES[:7]                      # the python record loop takes ~1 s per 100 k edges


@pytest.fixture(scope="session")
def sdbg_golden():
    with open(os.path.join(datasets.GOLDEN_DIR, "sdbg_golden.json")) as f:
        return json.load(f)


@pytest.mark.parametrize("ds,k,m", CPU_CASES)
@pytest.mark.parametrize("need_mult", [1, 0])
def test_oracle_index_matches_reference_dump(read_lib, sdbg_golden, ds, k, m, need_mult):
    _, rd = read_lib(ds)
    res = O.build_graph(rd, k, m)
    got = SO.digest(SO.build(res["stream"], np.asarray(res["meta"]), k, bool(need_mult)))
    exp = sdbg_golden["%s_k%d_m%d_mult%d" % (ds, k, m, need_mult)]
    assert set(got) == set(exp)
    bad = [s for s in exp if got[s] != exp[s]]
    assert not bad, bad


def device_sections(g, cabi, need_mult):
    """every array of a finished cabi.Sdbg under the section names of `sdbgdump`"""
    h = g.header()
    out = {"hdr": np.array([h.size, h.kmer_k, h.num_tips, h.words_per_tip_label], np.int64).tobytes(),
           "f": np.array(list(h.f), np.int64).tobytes(), "rank_f": np.array(list(h.rank_f), np.int64).tobytes(),
           "w_freq": np.array(list(h.w_freq), np.int64).tobytes(), "last_ones": np.int64(h.last_ones).tobytes(),
           "tip_ones": np.int64(h.tip_ones).tobytes()}
    for name, which in [("w", cabi.SDBG_W), ("last", cabi.SDBG_LAST), ("is_tip", cabi.SDBG_IS_TIP), ("invalid", cabi.SDBG_INVALID),
                        ("tip_seq", cabi.SDBG_TIP_SEQ), ("last_major", cabi.SDBG_LAST_MAJOR), ("last_minor", cabi.SDBG_LAST_MINOR),
                        ("last_sel", cabi.SDBG_LAST_SELECT), ("tip_major", cabi.SDBG_TIP_MAJOR), ("tip_minor", cabi.SDBG_TIP_MINOR)]:
        out[name] = g.array(which).tobytes()
    for c in range(9):
        out["w_major_%d" % c] = g.array(cabi.SDBG_W_MAJOR, c).tobytes()
        out["w_minor_%d" % c] = g.array(cabi.SDBG_W_MINOR, c).tobytes()
        out["w_sel_%d" % c] = g.array(cabi.SDBG_W_SELECT, c).tobytes()
    if need_mult:
        out["edge_multi"] = g.array(cabi.SDBG_EDGE_MULTI).tobytes()
        e = g.array(cabi.SDBG_LARGE_EDGE, 0, np.uint64).astype(np.int64)
        v = g.array(cabi.SDBG_LARGE_VALUE, 0, np.uint16).astype(np.int64)
        out["large_multi"] = np.stack([e, v], axis=1).tobytes() if len(e) else b""
    else:
        out["is_multi_1"] = g.array(cabi.SDBG_IS_MULTI_1).tobytes()
    return out


def build_on_device(cabi, rd, k, m, need_mult, how):
    with cabi.Context(k, m, device=0) as ctx:
        ctx.set_reads(rd["seq"], rd["start"], max_len=rd["max_len"])
        if m > 1:
            ctx.stage1()
        with cabi.Sdbg(k, bool(need_mult)) as g:
            if how == "device":                                   # records parsed where they lie in HBM
                g.from_stage2(ctx)
            else:                                                 # host deliveries, cut at arbitrary bucket boundaries
                stream, meta, _ = ctx.stage2()
                byts = meta[:, 0] * 2 + meta[:, 2] * 2 + meta[:, 1] * 4 * ((2 * k + 31) // 32)
                off = np.concatenate([[0], np.cumsum(byts)])
                cuts = [0, 1, 777, 16384, 40001, 65536] if how == "host_parts" else [0, 65536]
                for a, b in zip(cuts[:-1], cuts[1:]):
                    g.append(a, b, stream[int(off[a]):int(off[b])], meta[a:b])
            g.finish()
            return device_sections(g, cabi, need_mult)


@pytest.mark.gpu
@pytest.mark.parametrize("ds,k,m", CASES)
@pytest.mark.parametrize("need_mult,how", [(1, "device"), (0, "device"), (1, "host_parts"), (0, "host")])
def test_gpu_index_matches_reference(read_lib, sdbg_golden, ds, k, m, need_mult, how):
    from megagta_b200 import cabi
    _, rd = read_lib(ds)
    got = SO.digest(build_on_device(cabi, rd, k, m, need_mult, how))
    exp = sdbg_golden["%s_k%d_m%d_mult%d" % (ds, k, m, need_mult)]
    assert set(got) == set(exp)
    bad = [s for s in exp if got[s] != exp[s]]
    assert not bad, bad


@pytest.mark.gpu
def test_gpu_index_of_our_files_equals_the_reference_loader(read_lib, tmp_path):
    """megagta_b200 buildgraph writes the files, the UNMODIFIED SuccinctDBG::LoadFromMultiFile loads them (sdbgdump), the
    device builder gets the same records: every array equal, byte for byte (1 M reads: 27 k buckets in use, tips, large
    multiplicities, several major intervals)."""
    from megagta_b200 import cabi, sdbg_io
    if not O.have_ref() or "sdbgdump" not in open(O.REF_BIN, "rb").read().decode("latin1"):
        pytest.skip("oracle/_ref/megagta_ref with sdbgdump not built")
    prefix, rd = read_lib("meta1m")
    k, m = 31, 2
    ours = str(tmp_path / "ours")
    binp = os.path.join(os.path.dirname(cabi.LIB_PATH), "..", "bin", "megagta_b200")
    r = subprocess.run([binp, "buildgraph", "-k", str(k), "-m", str(m), "--host_mem", "4e9", "--num_cpu_threads", "4", "--num_output_threads", "1",
                        "--read_lib_file", prefix, "--output_prefix", ours], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    for need_mult in (1, 0):
        ref = SO.ref_dump(O.REF_BIN, ours, need_mult, str(tmp_path / "dump"))
        ref.pop("_load_seconds", None)
        got = build_on_device(cabi, rd, k, m, need_mult, "device")
        assert set(got) == set(ref)
        bad = [s for s in ref if got[s] != ref[s]]
        assert not bad, bad
        with cabi.Sdbg(k, bool(need_mult)) as g:                 # and straight from the files, like LoadFromMultiFile
            g.from_files(ours)
            g.finish()
            got = device_sections(g, cabi, need_mult)
        bad = [s for s in ref if got[s] != ref[s]]
        assert not bad, bad


@pytest.mark.gpu
def test_gpu_index_from_the_files_of_two_gpus(read_lib, tmp_path):
    """one .sdbg.<g> per GPU (MGTA_NUM_GPUS=2): the loader walks the sdbg_info rows across both files"""
    import torch
    from megagta_b200 import cabi
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    if not O.have_ref() or "sdbgdump" not in open(O.REF_BIN, "rb").read().decode("latin1"):
        pytest.skip("oracle/_ref/megagta_ref with sdbgdump not built")
    prefix, _ = read_lib("meta200k")
    ours = str(tmp_path / "ours")
    binp = os.path.join(os.path.dirname(cabi.LIB_PATH), "..", "bin", "megagta_b200")
    r = subprocess.run([binp, "buildgraph", "-k", "31", "-m", "2", "--host_mem", "4e9", "--num_cpu_threads", "4", "--num_output_threads", "1",
                        "--read_lib_file", prefix, "--output_prefix", ours], capture_output=True, text=True, timeout=900,
                       env=dict(os.environ, MGTA_NUM_GPUS="2"))
    assert r.returncode == 0, r.stderr[-2000:]
    ref = SO.ref_dump(O.REF_BIN, ours, 1, str(tmp_path / "dump"))
    ref.pop("_load_seconds", None)
    with cabi.Sdbg(31, True) as g:
        assert g.from_files(ours)["num_threads"] == 2
        g.finish()
        got = device_sections(g, cabi, 1)
    bad = [s for s in ref if got[s] != ref[s]]
    assert not bad, bad
