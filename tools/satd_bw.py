#!/usr/bin/env python3
"""SATD-8x8 whole-frame kernel throughput for each kernel form (X265CU_PIXELCMP_WIDE=0/1/2) against the HBM roofline.
usage: python tools/satd_bw.py [trace ...]"""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from harness import replay  # noqa: E402
from oracle import pyoracle as po  # noqa: E402
from src.x265_b200 import abi  # noqa: E402

peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0)
for name in sys.argv[1:] or ["c1_1080p", "c2_4k", "c3_4k10"]:
    trace = po.Trace(replay.trace_path(name))
    cfg = trace.cfg
    clip = replay.Clip(cfg)
    torch.cuda.set_device(0)
    r = bench.Runner(trace, clip, None, 0, True, torch)
    r.step()
    n = cfg["nframes"]
    P = 1 if cfg["depth"] == 8 else 2
    slots = np.array([abi.lib_host().x265cuh_frame_slot(r.frames[t]) for t in range(n)], np.int32)
    a, b = np.ascontiguousarray(np.tile(slots[:-1], 8)), np.ascontiguousarray(np.tile(slots[1:], 8))
    pix = len(a) * r.la.nCU * 64
    roof = peak / (2 * P + 4.0 / 64)
    for wide in ("0", "1", "2"):
        os.environ["X265CU_PIXELCMP_WIDE"] = wide
        ms = C.c_float()
        best = None
        for _ in range(6):
            abi.lib_cu().x265cu_pixelcmp_frames(r.la.ctx, 1, len(a), a.ctypes.data, b.ctypes.data, None, C.byref(ms))
            best = ms.value if best is None else min(best, ms.value)
        g = pix / (best * 1e-3) / 1e9
        print("%-9s wide=%s  %.3f ms  %.0f Gpix/s  %.2f of the HBM roofline (%.0f Gpix/s); distinct planes %d MB" % (
            name, wide, best, g, g / roof, roof, n * r.la.nCU * 64 * P >> 20))
    r.close()
