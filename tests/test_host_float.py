"""The float arithmetic of the host layer that must reproduce the reference's object code ("pinned compiler"):
cuTreeFinish's log2 mapping (encoder/slicetype.cpp:1844-1862).  The reference is built with g++ -O3 -ffast-math, which
evaluates `X265_LOG2(a) - X265_LOG2(b) + weightdelta` left to right with the libm log2; the host layer is compiled with the
same flags and keeps that order explicitly.  This test runs on the CPU (the mapping needs no GPU) and fails loudly when
another compiler or another flag set re-associates the expression or swaps the log2: element-wise bit equality against the
same expression evaluated in IEEE order with libm's log2 (CPython's math.log2), for the first-use path and for the memoised
path.  (The oracle's restatement of the same function is pinned against the reference traces' qpCuTreeOffset CRCs in
tests/test_oracle_golden.py.)"""
import ctypes as C
import math
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _inputs(n, seed):
    rng = np.random.default_rng(seed)
    intra = rng.integers(0, 6000, n, dtype=np.int32)
    intra[rng.integers(0, n, n // 50)] = 0                     # intracost == 0: the entry is left untouched
    invq = rng.integers(64, 1024, n, dtype=np.int32)
    prop = rng.integers(0, 65536, n).astype(np.uint16)
    prop[rng.integers(0, n, n // 10)] = 65535
    aq = rng.standard_normal(n) * 1.5
    return intra, invq, prop, aq


def _expected(intra, invq, prop, aq, fps_factor, weightdelta, strength, out):
    for i in range(len(intra)):
        intracost = (int(intra[i]) * int(invq[i]) + 128) >> 8
        if intracost:
            p = (int(prop[i]) * fps_factor + 128) >> 8
            diff = math.log2(float(intracost + p)) - math.log2(float(intracost))
            out[i] = float(aq[i]) - strength * (diff + weightdelta)
    return out


def test_cutree_finish_mapping_bit_exact(built):
    from src.x265_b200 import abi
    L = abi.lib_host()
    fn = L.x265cuh_cutree_finish_map
    fn.restype = None
    fn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double]
    n = 8160
    for seed, fps_factor, weightdelta, strength in ((1, 256, 0.0, 2.0), (2, 256, 1.0 - 0.9731, 2.0), (3, 213, 0.25, 5.0 * (1.0 - 0.7)), (4, 307, 0.0, 1.5)):
        intra, invq, prop, aq = _inputs(n, seed)
        sentinel = np.full(n, 123.456)
        want = _expected(intra, invq, prop, aq, fps_factor, weightdelta, strength, sentinel.copy())
        for _ in range(2):                                        # first use fills the memo table, the second reads it
            got = sentinel.copy()
            fn(intra.ctypes.data, invq.ctypes.data, prop.ctypes.data, aq.ctypes.data, got.ctypes.data, n, fps_factor, weightdelta, strength)
            bad = np.nonzero(got.view(np.uint64) != want.view(np.uint64))[0]
            assert bad.size == 0, "qpCuTreeOffset differs from the reference's evaluation order at %d of %d CUs (first %d: %r vs %r): the host " \
                                  "layer must be compiled with g++ -O3 -ffast-math as the reference is (DESIGN.md 5)" % (
                                      bad.size, n, bad[0], got[bad[0]], want[bad[0]])
