#!/usr/bin/env python3
"""Static SASS statistics per kernel of libx265cu.so (instruction count, opcode mix).
usage: tools/sass_stats.py [kernel-substring]   -- runs cuobjdump -sass on the built library"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "src", "x265_b200", "libx265cu.so")
want = sys.argv[1] if len(sys.argv) > 1 else ""
out = subprocess.run(["cuobjdump", "-sass", lib], stdout=subprocess.PIPE, text=True).stdout
cur, stats = None, {}
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        stats[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        stats[cur][m.group(2)] += 1
for k, c in stats.items():
    if want and want not in k:
        continue
    name = subprocess.run(["c++filt", k], stdout=subprocess.PIPE, text=True).stdout.strip()[:90]
    print("%-90s %6d instr" % (name, sum(c.values())))
    if want:
        print("   ", ", ".join("%s %d" % kv for kv in c.most_common(24)))
