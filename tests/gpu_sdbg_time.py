"""Timing of the device SdBG load + rank/select build against the reference's SuccinctDBG::LoadFromMultiFile on the same
graph (not a pytest file).  usage: gpu_sdbg_time.py N_READS [k] [m]   -> one JSON line"""
import json, os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from megagta_b200 import cabi, synth
from oracle import oracle as O
from oracle import sdbg_oracle as SO

n = int(sys.argv[1]); k = int(sys.argv[2]) if len(sys.argv) > 2 else 31; m = int(sys.argv[3]) if len(sys.argv) > 3 else 2
work = tempfile.mkdtemp(prefix="mgta_sdbg_time_")
prefix = os.path.join(work, "reads")
synth.write_metagenome(prefix, n, 150)
rd = O.load_read_lib(prefix)
out = {"reads": n, "k": k, "m": m}
with cabi.Context(k, m) as ctx:
    ctx.set_reads(rd["seq"], rd["start"], max_len=rd["max_len"])
    ctx.stage1()
    for rep in range(3):                                            # stage 2 with the records parsed in HBM + tables
        with cabi.Sdbg(k, True) as g:
            t = time.time(); g.from_stage2(ctx); t1 = time.time() - t
            t = time.time(); h = g.finish(); t2 = time.time() - t
            out["edges"] = int(h.size); out["device_stage2_into_sdbg_s"] = t1; out["device_tables_s"] = t2
            out["stage2_ms_device"] = ctx.stats(2)["ms_total"]
    t = time.time(); ctx.stage2(collect=False); out["stage2_alone_s"] = time.time() - t
# the reference: its own loader on files written by our driver
ours = os.path.join(work, "ours")
binp = os.path.join(ROOT, "megagta_b200", "bin", "megagta_b200")
r = subprocess.run([binp, "buildgraph", "-k", str(k), "-m", str(m), "--host_mem", "64e9", "--num_cpu_threads", "8", "--num_output_threads", "2",
                    "--read_lib_file", prefix, "--output_prefix", ours], capture_output=True, text=True)
assert r.returncode == 0, r.stderr[-2000:]
if O.have_ref():
    t = time.time()
    d = SO.ref_dump(O.REF_BIN, ours, 1, os.path.join(work, "dump"))
    out["reference_load_and_dump_s"] = time.time() - t
    out["reference_edges"] = int(np.frombuffer(d["hdr"], np.int64)[0])
    out["reference_load_s"] = float(d.get("_load_seconds", b"nan"))
print(json.dumps(out))
