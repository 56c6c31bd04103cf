"""SURVEY 8f-4, second half: MotionEstimate::motionEstimate on full-resolution planes (encoder/motion.cpp:571-1172) as
x265cu_motion_estimate -- every integer pattern (DIA, HEX, UMH, STAR, FULL), every sub-pel level, the 24 inter PU shapes.
The CUDA path == the committed results of the unmodified reference (tests/golden/me_golden.json) == the oracle on larger
seeded batches (the oracle itself is pinned against the reference in tests/test_oracle_vs_ref.py and tests/test_me_golden.py)."""
import ctypes as C
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "me_golden.json")


@pytest.fixture(scope="module")
def me(built):
    from harness import me_cases as mc
    from oracle import pyoracle as po
    from src.x265_b200 import abi
    oracle_me, oracle_run = mc.oracle_me, mc.oracle_run
    L = abi.lib_cu()
    if L.x265cu_device_count() < 1:
        pytest.fail("no CUDA device visible: the GPU tests must run on the B200 box (there is no CPU fallback)")
    V, I, S, Z = C.c_void_p, C.c_int, C.c_ssize_t, C.c_size_t
    L.x265cu_motion_estimate.argtypes = [V, I, I, I, I, V, Z, S, V, Z, S, V, I, V, V, V]
    return mc, po, abi, L, oracle_me, oracle_run


def gpu_run(me, la, case, lut, built):
    mc, po, abi, L, _, _ = me
    fenc, ref, stride, its = built
    n = case["n"]
    out = (mc.MeResult * n)()
    r = L.x265cu_motion_estimate(la.ctx, mc.METHODS[case["method"]], case["subme"], case["w"], case["h"], fenc.ctypes.data, fenc.size, stride,
                                 ref.ctypes.data, ref.size, stride, lut.ctypes.data + 65536 * 2, n, C.addressof(its), C.addressof(out), None)
    assert r == 0, L.x265cu_last_error(la.ctx)
    return mc.results_list(out, n)


@pytest.mark.parametrize("depth", [8, 10])
def test_motion_search_equals_reference_results(me, depth):
    mc, po, abi, L, oracle_me, oracle_run = me
    gold = json.load(open(GOLDEN))["depths"][str(depth)]
    cases = mc.golden_cases()
    lut = mc.make_lut(20)
    la = abi.Lookahead(64, 64, depth, 1, 4, 0, 0, 0, 0, 0.0)
    try:
        for case, want in zip(cases, gold):
            got = gpu_run(me, la, case, lut, mc.build_case(case, depth))
            want = [tuple(r) for r in want]
            assert got == want, (case, [(i, a, b) for i, (a, b) in enumerate(zip(got, want)) if a != b][:3])
    finally:
        la.close()


@pytest.mark.parametrize("depth", [8, 10])
@pytest.mark.parametrize("method", ["dia", "hex", "umh", "star", "full"])
def test_motion_search_equals_oracle(me, depth, method):
    """larger seeded batches than the goldens hold: every sub-pel level, a square, a rectangular and an AMP shape, other
    mvcost tables (lambda 1.25 and 40), a picture where half of the windows are cut"""
    mc, po, abi, L, oracle_me, oracle_run = me
    O = oracle_me(depth)
    la = abi.Lookahead(64, 64, depth, 1, 4, 0, 0, 0, 0, 0.0)
    try:
        seed = 5000 + 100 * mc.METHODS[method]
        for subme in range(8):
            for (w, h), lam in (((16, 16), 5), ((32, 8), 160), ((12, 16), 20), ((64, 64), 20)):
                if (w, h) == (64, 64) and (subme not in (2, 7) or method == "full"):
                    continue
                seed += 1
                n = 8 if method == "full" else (24 if w == 64 else 96)
                kw = dict(merange_choices=(4, 9)) if method == "full" else dict(tight=0.5)
                case = dict(method=method, subme=subme, w=w, h=h, n=n, seed=seed, W=224, H=160, kw=kw)
                built = mc.build_case(case, depth)
                lut = mc.make_lut(lam)
                got = gpu_run(me, la, case, lut, built)
                want = oracle_run(O, case, depth, lut, built)
                assert got == want, (case, [(i, a, b) for i, (a, b) in enumerate(zip(got, want)) if a != b][:3])
    finally:
        la.close()


def test_motion_search_refuses_bad_input(me):
    mc, po, abi, L, _, _ = me
    la = abi.Lookahead(64, 64, 8, 1, 4, 0, 0, 0, 0, 0.0)
    try:
        case = dict(method="star", subme=2, w=16, h=16, n=4, seed=1, W=64, H=64, kw={})
        fenc, ref, stride, its = mc.build_case(case, 8)
        lut = mc.make_lut(20)
        out = (mc.MeResult * 4)()

        def call(w=16, h=16, method=3, subme=2, samples=None, n=4):
            return L.x265cu_motion_estimate(la.ctx, method, subme, w, h, fenc.ctypes.data, fenc.size, stride, ref.ctypes.data, samples or ref.size, stride,
                                            lut.ctypes.data + 65536 * 2, n, C.addressof(its), C.addressof(out), None)
        assert call() == 0
        assert call(w=4, h=4) == -1            # not an inter PU (motion.cpp:168)
        assert call(w=24, h=24) == -1
        assert call(method=5) == -1 and call(subme=8) == -1
        assert call(samples=ref.size // 2) == -1     # a window that leaves the plane is refused, not clamped
        its[2].merange = 0
        assert call() == -1
        its[2].merange = 16
        its[1].numCandidates = 13
        assert call() == -1
        assert call(n=0) == 0
    finally:
        la.close()
