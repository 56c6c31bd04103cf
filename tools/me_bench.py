"""Throughput of the full-resolution PU motion search on one GPU against the oracle port on one host core:
every aligned 16x16 PU of a 1080p picture, STAR / UMH / HEX, sub-pel level 2, merange 57, predictors = zero vector.
Prints searches/s (kernel time from CUDA events; the call with host planes in brackets)."""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from harness import me_cases as mc        # noqa: E402
from oracle import pyoracle as po         # noqa: E402
from src.x265_b200 import abi             # noqa: E402


def main():
    depth = int(os.environ.get("DEPTH", "8"))
    W, H, w, h = 1920, 1080, int(os.environ.get("PU", "16")), int(os.environ.get("PU", "16"))
    fenc, ref, stride, origin = mc.planes(depth, W, H, 9, motion=(7, -4), noise=3)
    xs, ys = np.meshgrid(np.arange(0, W - w + 1, w), np.arange(0, H - h + 1, h))
    n = xs.size
    its = (mc.MeItem * n)()
    for i, (x, y) in enumerate(zip(xs.ravel(), ys.ravel())):
        it = its[i]
        it.offset = origin + int(y) * stride + int(x)
        it.merange = 57
        it.mvmin[0], it.mvmin[1] = max(-57, -int(x) - 24), max(-57, -int(y) - 24)
        it.mvmax[0], it.mvmax[1] = min(57, W - int(x) - w + 24), min(57, H - int(y) - h + 24)
    lut = mc.make_lut(20)
    L = abi.lib_cu()
    V, I, S, Z = C.c_void_p, C.c_int, C.c_ssize_t, C.c_size_t
    L.x265cu_motion_estimate.argtypes = [V, I, I, I, I, V, Z, S, V, Z, S, V, I, V, V, V]
    la = abi.Lookahead(64, 64, depth, 1, 4, 0, 0, 0, 0, 0.0)
    O = mc.oracle_me(depth)
    out = (mc.MeResult * n)()
    for name in ("star", "umh", "hex"):
        for subme in (2, 5):
            ms = C.c_float()
            best, wall = 1e9, 1e9
            for rep in range(4):
                t0 = time.perf_counter()
                r = L.x265cu_motion_estimate(la.ctx, mc.METHODS[name], subme, w, h, fenc.ctypes.data, fenc.size, stride, ref.ctypes.data, ref.size, stride,
                                             lut.ctypes.data + 65536 * 2, n, C.addressof(its), C.addressof(out), C.byref(ms))
                assert r == 0
                wall = min(wall, time.perf_counter() - t0)
                best = min(best, ms.value)
            m = min(n, 600)
            want = (mc.MeResult * m)()
            t0 = time.perf_counter()
            O.ola_motion_estimate_batch(mc.METHODS[name], subme, w, h, fenc.ctypes.data, stride, ref.ctypes.data, stride, lut.ctypes.data + 65536 * 2, m,
                                        C.addressof(its), C.addressof(want))
            cpu = time.perf_counter() - t0
            ok = mc.results_list(want, m) == mc.results_list(out, m)
            print("%-4s subme %d  %dx%d x %d: kernel %.3f ms = %.2f M searches/s (call with host planes %.2f ms); oracle port, 1 core: %.1f k searches/s; parity %s"
                  % (name, subme, w, h, n, best, n / best / 1e3, wall * 1e3, m / cpu / 1e3, "ok" if ok else "MISMATCH"))
    la.close()


if __name__ == "__main__":
    main()
