#!/usr/bin/env python3
"""Host wall-clock split of one replay step by call type (pre-lookahead frames, batches, single estimates),
in host-buffer (e2e) and device-resident mode.  usage: python tools/e2e_profile.py [trace]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from harness import replay  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c1_1080p"
trace = po.Trace(replay.trace_path(name))
clip = replay.Clip(trace.cfg)
torch.cuda.set_device(0)
for resident in (True, False):
    r = bench.Runner(trace, clip, None, 0, resident, torch)
    for _ in range(2):
        r.step()
    acc = {"P": [0.0, 0], "B": [0.0, 0], "J": [0.0, 0], "T": [0.0, 0], "sync": [0.0, 0]}
    t_all = time.perf_counter()
    for c in r.calls:
        t0 = time.perf_counter()
        if c[0] == "P":
            r.la.pre_lookahead_batch_prepared(c[2], os.environ.get('PLANES_BACK', '1') == '1')
            k = "P"
        elif c[0] == "T":
            r.la.cutree_sequence_prepared(c[1])
            k = "T"
        else:
            r.la.estimate_prepared(c[1], c[2])
            k = "B" if c[2] else "J"
        acc[k][0] += time.perf_counter() - t0
        acc[k][1] += 1
    t0 = time.perf_counter()
    r.la.sync()
    acc["sync"][0] += time.perf_counter() - t0
    acc["sync"][1] += 1
    total = time.perf_counter() - t_all
    print("resident" if resident else "host-buffers", "total %.1f ms:" % (total * 1e3),
          ", ".join("%s %.1f ms / %d calls" % (k, v[0] * 1e3, v[1]) for k, v in acc.items()))
    for _ in range(int(os.environ.get('EXTRA_STEPS', '0'))):
        r.step()
    r.close()
