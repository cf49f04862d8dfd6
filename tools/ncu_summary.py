#!/usr/bin/env python
"""ncu .ncu-rep -> markdown summary for profiles/ (run in the build container: `ncu -i` needs no GPU).
usage: tools/ncu_summary.py REPORT.ncu-rep "title line" > profiles/NAME.md"""
import csv
import io
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_shared_mem",
           "launch__occupancy_limit_registers", "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
           "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct"]
STALL = "smsp__average_warps_issue_stalled_"


def main():
    rep, title = sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    print("# " + title + "\n")
    print("Source report: %s (scratch, not committed).  traffic = dram__bytes_read.sum + dram__bytes_write.sum per launch.\n" % rep)
    for r in data:
        if float(r[idx["gpu__time_duration.sum"]]) < 0.05 and units[idx["gpu__time_duration.sum"]] == "ms":
            continue
        print("## " + r[idx["Kernel Name"]] + "\n")
        for m in METRICS:
            if m in idx:
                print("* %s = %s %s" % (m, r[idx[m]], units[idx[m]]))
        try:
            t = float(r[idx["dram__bytes_read.sum"]]) + float(r[idx["dram__bytes_write.sum"]])
            ms = float(r[idx["gpu__time_duration.sum"]])
            print("* traffic = %.3f %s -> %.0f GB/s" % (t, units[idx["dram__bytes_read.sum"]], t / ms * 1000.0))
        except Exception:
            pass
        st = sorted(((float(r[i]), h[len(STALL):].replace("_per_issue_active.ratio", "")) for h, i in idx.items()
                     if h.startswith(STALL) and h.endswith("_per_issue_active.ratio") and r[i]), reverse=True)[:5]
        print("* top stalls (warps per issue): " + ", ".join("%s %.1f" % (n, v) for v, n in st) + "\n")


if __name__ == "__main__":
    main()
