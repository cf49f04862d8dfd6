"""Real multi-GPU parity (one process per GPU, NCCL): skipped on boxes with a single GPU.  The single-device tests
tests/test_gpu_sharded.py walks the same protocol with all shards on one device."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_two_gpus_reproduce_the_reference_goldens(data_dir):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (have %d)" % n)
    env = dict(os.environ, MGTA_TEST_DATA=data_dir)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
                        os.path.join(ROOT, "tests", "gpu_multi.py")], capture_output=True, text=True, timeout=900, env=env)
    sys.stdout.write(r.stdout[-4000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "MISMATCH" not in r.stdout


def test_cpp_driver_on_two_gpus_writes_one_file_per_gpu(golden, read_lib, tmp_path):
    """`megagta_b200 buildgraph` with one host thread per GPU and NCCL between them (MGTA_NUM_GPUS=2): <p>.sdbg.0/.1, an
    sdbg_info with num_threads 2 whose rows name the file of every bucket (sdbg_multi_io.h:160-187), and the logical
    stream of the unmodified reference."""
    import torch
    from megagta_b200 import sdbg_io
    from oracle import oracle as O
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    binary = os.path.join(ROOT, "megagta_b200", "bin", "megagta_b200")
    for case in ("meta200k_k31_m2", "meta200k_k61_m2", "smoke_k31_m1"):
        g = golden["cases"][case]
        prefix, _ = read_lib(g["dataset"])
        out = str(tmp_path / case)
        r = subprocess.run([binary, "buildgraph", "-k", str(g["k"]), "-m", str(g["m"]), "--host_mem", "4e9", "--num_cpu_threads", "4",
                            "--num_output_threads", "1", "--read_lib_file", prefix, "--output_prefix", out],
                           capture_output=True, text=True, timeout=600, env=dict(os.environ, MGTA_NUM_GPUS="2"))
        assert r.returncode == 0, r.stderr[-3000:]
        hdr, stream, meta = sdbg_io.canonical(out)
        assert hdr["num_threads"] == 2 and os.path.exists(out + ".sdbg.1")
        _, rows = sdbg_io.read_info(out)
        assert set(rows[rows[:, 3] > 0][:, 1].tolist()) == {0, 1}            # both files hold buckets
        assert hdr["total_size"] == g["total_size"] and hdr["num_tips"] == g["num_tips"]
        assert O.stream_hash(stream) == g["stream_hash"] and O.meta_hash(meta) == g["meta_hash"]


def test_cpp_driver_with_mercy_on_two_gpus(golden, read_lib, tmp_path):
    """--need_mercy with MGTA_NUM_GPUS=2 (real NCCL: is_solid all-reduce, candidate all-gathers): the graph and "Number mercy"
    of the unmodified reference run with --need_mercy"""
    import re
    import torch
    from megagta_b200 import sdbg_io
    from oracle import oracle as O
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    binary = os.path.join(ROOT, "megagta_b200", "bin", "megagta_b200")
    for case in ("smoke_k31_m2_mercy", "adversarial_k27_m3_mercy"):
        g = golden["cases"][case]
        prefix, _ = read_lib(g["dataset"])
        out = str(tmp_path / case)
        r = subprocess.run([binary, "buildgraph", "-k", str(g["k"]), "-m", str(g["m"]), "--host_mem", "4e9", "--num_cpu_threads", "4",
                            "--num_output_threads", "1", "--read_lib_file", prefix, "--output_prefix", out, "--need_mercy"],
                           capture_output=True, text=True, timeout=600, env=dict(os.environ, MGTA_NUM_GPUS="2"))
        assert r.returncode == 0, r.stderr[-3000:]
        hdr, stream, meta = sdbg_io.canonical(out)
        assert hdr["num_threads"] == 2
        assert int(re.search(r"Number mercy: (\d+)", r.stderr).group(1)) == g["num_mercy"]
        assert O.stream_hash(stream) == g["stream_hash"] and O.meta_hash(meta) == g["meta_hash"]
