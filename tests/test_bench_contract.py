"""bench.py contract on the CPU: the reference arm (`--impl reference`) runs the reference's own CPU CX1 on a bounded
sample and prints one JSON line with the keys the driver reads; the GPU arm's line is checked on the GPU box."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

REQUIRED = ["impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
            "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"]


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--sample-reads", "30000",
                        "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for k in REQUIRED:
        assert k in line, k
    assert line["impl"] == "reference" and line["metric"] == "SdBG edges/sec" and line["unit"] == "edges/s"
    assert line["value"] > 0 and line["higher_is_better"] is True
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in line["config"] and "sample" in line["config"]
    assert line["reads_per_s"] > 0 and len(line["config"]["sample_hashes"]["stream_hash"]) == 16


@pytest.mark.gpu
def test_gpu_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--reads-per-gpu", "1000000", "--steps", "1", "--warmup", "3",
                        "--sample-reads", "20000"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for k in ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"]:
        assert k in line, k
    assert line["gpu_launches"] > 0 and line["value"] > 0 and line["e2e"]["value"] > 0
    assert line["e2e"]["h2d_bytes_per_step"] > 0 and line["e2e"]["d2h_bytes_per_step"] > 0
    for k in ["bound", "achieved", "peak", "unit", "frac", "traffic"]:
        assert k in line["roofline"], k
    for k in ["value", "unit", "cores", "kind", "sample"]:
        assert k in line["cpu_baseline"], k
    par = line["config"]["parity"]
    assert par["sample"]["ok"] is True, par["sample"]          # GPU path == reference binary on the shared sample, in this job
    assert len(par["stream_hash"]) == 16 and par["total_size"] == line["config"]["edges_per_step"]
