#!/usr/bin/env python3
"""Event-timed kernel milliseconds per family of one device-resident replay step (after warm-up).
usage: python tools/kernel_ms.py <workload> [steps]   (experiment switches through the environment)"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from harness import replay  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c2_4k"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
trace = po.Trace(replay.trace_path(name))
clip = replay.Clip(trace.cfg)
torch.cuda.set_device(0)
r = bench.Runner(trace, clip, None, 0, True, torch)
for _ in range(2):
    r.step()
r.la.stats_enable(True)
r.la.stats(reset=True)
for _ in range(steps):
    r.step()
st = r.la.stats(reset=True)
r.close()
print(json.dumps({"workload": name, "ms": {k: round(v / steps, 4) for k, v in st["ms"].items() if v}, "launches": {k: v // steps for k, v in st["launches"].items() if v}}))
