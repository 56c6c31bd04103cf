#!/usr/bin/env python3
"""Per-source-line instruction and stall-sample totals from an ncu report (needs -lineinfo + --import-source on).
usage: tools/ncu_lines.py <report.ncu-rep> [top N]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fname = ""
lines = []
hdr = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        ci = hdr.index("Instructions Executed")
        si = hdr.index("Warp Stall Sampling (All Samples)")
        continue
    if hdr and r[0].isdigit():
        try:
            lines.append((int(r[ci]), int(r[si]), fname, int(r[0]), r[1].strip()))
        except ValueError:
            pass
ti = sum(l[0] for l in lines)
ts = sum(l[1] for l in lines)
print("total instructions %d, stall samples %d" % (ti, ts))
for l in sorted(lines, key=lambda x: -x[1])[:top]:
    print("%5.1f%% inst %5.1f%% stall  %s:%d  %s" % (100.0 * l[0] / max(1, ti), 100.0 * l[1] / max(1, ts), l[2], l[3], l[4][:110]))
