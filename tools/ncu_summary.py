#!/usr/bin/env python3
"""Summarise an ncu report (raw metrics + SASS-level stall/opcodes) for one kernel launch.
usage: tools/ncu_summary.py <report.ncu-rep> [launch index]"""
import collections
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
idx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_executed.avg.per_cycle_elapsed", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_miss.sum",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]
r = data[idx]
for w in want:
    if w in hdr:
        i = hdr.index(w)
        print("%-70s %-10s %s" % (w, units[i], r[i][:90]))
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], stdout=subprocess.PIPE, text=True).stdout
srows = list(csv.reader(io.StringIO(sass)))
secs = [i for i, x in enumerate(srows) if x and x[0] == "Address"]
h = srows[secs[idx]]
end = secs[idx + 1] - 1 if idx + 1 < len(secs) else len(srows)
body = srows[secs[idx] + 1:end]
ci, si = h.index("Instructions Executed"), h.index("Warp Stall Sampling (All Samples)")
tot = sum(int(x[ci]) for x in body if x[ci].isdigit())
print("instructions executed: %d  (static %d)" % (tot, len(body)))
st = {}
for i, n in enumerate(h):
    if n.startswith("stall_") and "Not Issued" not in n:
        st[n] = sum(int(x[i]) for x in body if len(x) > i and x[i].isdigit())
ssum = max(1, sum(st.values()))
print("stall samples:", ", ".join("%s %.1f%%" % (n[6:], 100.0 * v / ssum) for n, v in sorted(st.items(), key=lambda kv: -kv[1])[:9]))
op, ops = collections.Counter(), collections.Counter()
for x in body:
    if not x[ci].isdigit():
        continue
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", x[1])
    if m:
        op[m.group(2)] += int(x[ci])
        ops[m.group(2)] += int(x[si]) if x[si].isdigit() else 0
print("opcodes:", ", ".join("%s %.1f%%/%d" % (o, 100.0 * v / tot, ops[o]) for o, v in op.most_common(16)))
