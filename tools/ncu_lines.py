#!/usr/bin/env python
"""Per-source-line instruction / stall-sample summary of one kernel from an .ncu-rep captured with --import-source on
(kernels compiled with -lineinfo).  usage: tools/ncu_lines.py REPORT.ncu-rep KERNEL_REGEX [min_pct]   (`ncu -i`, no GPU)"""
import csv
import io
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    min_pct = float(sys.argv[3]) if len(sys.argv) > 3 else 0.7
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "-k", "regex:" + kern],
                         capture_output=True, text=True).stdout
    fname, hdr, lines = None, None, []
    for r in csv.reader(io.StringIO(raw)):
        if len(r) >= 2 and r[0] == "File Path":
            fname = r[1].split("/")[-1]; hdr = None
        elif r and r[0] == "Line No":
            hdr = {}
            for i, h in enumerate(r):
                hdr.setdefault(h, i)
        elif hdr and r and r[0].isdigit():
            def num(h):
                v = r[hdr[h]]
                try:
                    return float(v)
                except ValueError:
                    return 0.0
            lines.append((fname, int(r[0]), r[1], num("Instructions Executed"), num("# Samples"), num("Avg. Threads Executed")))
    ti = sum(x[3] for x in lines) or 1.0
    ts = sum(x[4] for x in lines) or 1.0
    print("total warp instructions %.3e, stall samples %d" % (ti, ts))
    for f, ln, src, inst, samp, thr in lines:
        if inst / ti * 100 >= min_pct or samp / ts * 100 >= min_pct:
            print("%-16s %4d inst %5.1f%% samp %5.1f%% thr %4.1f | %s" % (f, ln, inst / ti * 100, samp / ts * 100, thr, src.strip()[:105]))


if __name__ == "__main__":
    main()
