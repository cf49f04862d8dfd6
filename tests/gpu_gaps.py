"""Where a device-only step spends its time (not a pytest file): per step, wall time of each C call and of the Python in
between, the stages' own device windows (mgta_get_stats ms_total) and the sum of their timed kernels.
usage: gpu_gaps.py [N_READS] [k] [steps]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from megagta_b200 import cabi, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 31
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 8
seq, start = synth.packed_metagenome(n, 150, procs=16)
stream = torch.cuda.Stream()
with torch.cuda.stream(stream), cabi.Context(k, 2, stream=stream.cuda_stream) as ctx:
    ctx.set_reads(seq, start, max_len=150)
    for i in range(steps + 3):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(stream)
        ctx.stage1()
        t1 = time.perf_counter()
        s1 = ctx.stats(1)
        t2 = time.perf_counter()
        ctx.stage2(collect=False)
        t3 = time.perf_counter()
        e1.record(stream)
        torch.cuda.synchronize()
        s2 = ctx.stats(2)
        ph = lambda s: s["ms_extract"] + s["ms_partition"] + s["ms_sort_emit"] + s["ms_hist"] + s.get("ms_nodes", 0.0)
        if i >= 3:
            print("step %d: events %.1f ms | wall stage1 %.1f  between %.2f  stage2 %.1f | windows s1 %.1f (kernels %.1f) s2 %.1f (kernels %.1f) nodes %.1f"
                  % (i - 3, e0.elapsed_time(e1), (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, s1["ms_total"], ph(s1), s2["ms_total"], ph(s2),
                     s2.get("ms_nodes", 0.0)), flush=True)
    print("s1", ctx.stats(1))
    print("s2", ctx.stats(2))
