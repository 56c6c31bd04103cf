"""Pin of the oracle, primitive by primitive, against the reference's OWN compiled C functions
(setupCPrimitives table exposed by oracle/_ref/libx265ref<depth>.so, built from the unmodified
sources by oracle/build_ref.py).  Input recipes follow the reference's TestBench
(test/pixelharness.cpp:30-62,80-200,318-359,499-527; test/intrapredharness.cpp:47-140).

Runs only where oracle/_ref exists (the build container); the committed golden traces carry the
same pin to machines without the reference (tests/test_oracle_golden.py)."""
import ctypes as C

import numpy as np
import pytest

from oracle import pyoracle as po

DEPTHS = [d for d in (8, 10) if po.ref_available(d)]
pytestmark = pytest.mark.skipif(not DEPTHS, reason="oracle/_ref not built (reference tree absent)")


def bufs(depth, seed, n=64 * 64 * 2):
    dt = po.pixel_dtype(depth)
    rng = np.random.default_rng(seed)
    mx = (1 << depth) - 1
    return [rng.integers(0, mx + 1, n).astype(dt), np.zeros(n, dt), np.full(n, mx, dt),
            rng.integers(0, 4, n).astype(dt), (mx - rng.integers(0, 4, n)).astype(dt)]


def ptr(a, off=0):
    return a.ctypes.data + off * a.itemsize


@pytest.mark.parametrize("depth", DEPTHS)
@pytest.mark.parametrize("name", ["sad8x8", "satd8x8", "sa8d8x8", "sa8d16x16"])
def test_pixelcmp(depth, name):
    O, R = po.oracle(depth), po.ref(depth)
    fo, fr = getattr(O, "ola_" + name), getattr(R, "x265ref_" + name)
    B = bufs(depth, 11)
    for a in B:
        for b in B:
            for it in range(100):
                oa, ob = (it * 32) % 3000, (it * 37 + 5) % 3000
                assert fo(ptr(a, oa), 64, ptr(b, ob), 59) == fr(ptr(a, oa), 64, ptr(b, ob), 59)


@pytest.mark.parametrize("depth", DEPTHS)
def test_sad_x3_x4(depth):
    O, R = po.oracle(depth), po.ref(depth)
    B = bufs(depth, 12)
    fenc = np.zeros(64 * 8, B[0].dtype)   # FENC_STRIDE 64
    res = (C.c_int32 * 4)()
    for a in B[:3]:
        for b in B[:3]:
            fenc[:] = a[:64 * 8]
            for it in range(50):
                offs = [(it * 7 + k * 13) % 2000 for k in range(4)]
                R.x265ref_sad_x4_8x8(ptr(fenc), ptr(b, offs[0]), ptr(b, offs[1]), ptr(b, offs[2]), ptr(b, offs[3]), 59, res)
                want = [O.ola_sad8x8(ptr(fenc), 64, ptr(b, o), 59) for o in offs]
                assert list(res) == want
                R.x265ref_sad_x3_8x8(ptr(fenc), ptr(b, offs[0]), ptr(b, offs[1]), ptr(b, offs[2]), 59, res)
                assert list(res)[:3] == want[:3]


@pytest.mark.parametrize("depth", DEPTHS)
def test_pixelavg(depth):
    O, R = po.oracle(depth), po.ref(depth)
    B = bufs(depth, 13)
    for a in B:
        for b in B:
            d1 = np.zeros(64 * 8, a.dtype)
            d2 = np.zeros(64 * 8, a.dtype)
            O.ola_pixelavg8x8(ptr(d1), 64, ptr(a, 17), 61, ptr(b, 29), 67)
            R.x265ref_pixelavg8x8(ptr(d2), 64, ptr(a, 17), 61, ptr(b, 29), 67)
            assert np.array_equal(d1, d2)


@pytest.mark.parametrize("depth", DEPTHS)
def test_frame_init_lowres_and_border(depth):
    O, R = po.oracle(depth), po.ref(depth)
    dt = po.pixel_dtype(depth)
    rng = np.random.default_rng(14)
    w, h, mx, my = 40, 24, 48, 32
    sstride = 2 * w + 17
    src = rng.integers(0, 1 << depth, (2 * h + 1) * sstride).astype(dt)
    dstride = w + 2 * mx + 8
    size = dstride * (h + 2 * my)
    outs = []
    for L, pre in ((O, "ola_"), (R, "x265ref_")):
        planes = [np.zeros(size, dt) for _ in range(4)]
        base = dstride * my + mx
        getattr(L, pre + "frame_init_lowres")(ptr(src), *[ptr(p, base) for p in planes], sstride, dstride, w, h)
        for p in planes:
            getattr(L, pre + "extend_border")(ptr(p, base), dstride, w, h, mx, my)
        outs.append(np.concatenate(planes))
    assert np.array_equal(outs[0], outs[1])


@pytest.mark.parametrize("depth", DEPTHS)
def test_intra_filter_and_predictors(depth):
    O, R = po.oracle(depth), po.ref(depth)
    dt = po.pixel_dtype(depth)
    rng = np.random.default_rng(15)
    mx = (1 << depth) - 1
    for it in range(100):
        kind = it % 4
        if kind == 0:
            nb = rng.integers(0, mx + 1, 33).astype(dt)
        elif kind == 1:
            nb = np.zeros(33, dt)
        elif kind == 2:
            nb = np.full(33, mx, dt)
        else:
            nb = np.where(rng.integers(0, 2, 33) > 0, mx, 0).astype(dt)
        f1, f2 = np.zeros(33, dt), np.zeros(33, dt)
        O.ola_intra_filter8(ptr(nb), ptr(f1))
        R.x265ref_intra_filter8(ptr(nb), ptr(f2))
        assert np.array_equal(f1, f2)
        for mode in range(35):
            for bf in (0, 1):
                p1, p2 = np.zeros(64, dt), np.zeros(64, dt)
                O.ola_intra_pred8(mode, ptr(p1), 8, ptr(nb), bf)
                R.x265ref_intra_pred8(mode, ptr(p2), 8, ptr(nb), bf)
                assert np.array_equal(p1, p2), (mode, bf)


@pytest.mark.parametrize("depth", DEPTHS)
def test_weight_pp_and_var(depth):
    O, R = po.oracle(depth), po.ref(depth)
    B = bufs(depth, 16)
    corr = 14 - depth
    for a in B:
        for (w0, denom, off) in ((64, 6, 0), (70, 6, -3), (127, 7, 12), (33, 5, -128), (1, 0, 127), (90, 7, 5)):
            rnd = (1 << (denom - 1)) if denom else 0
            d1, d2 = np.zeros(64 * 16, a.dtype), np.zeros(64 * 16, a.dtype)
            O.ola_weight_pp(ptr(a), ptr(d1), 64, 64, 16, w0, rnd << corr, denom + corr, off << (depth - 8))
            R.x265ref_weight_pp(ptr(a), ptr(d2), 64, 64, 16, w0, rnd << corr, denom + corr, off << (depth - 8))
            assert np.array_equal(d1, d2)
        for o in (0, 5, 77):
            assert O.ola_var16(ptr(a, o), 64) == R.x265ref_var16(ptr(a, o), 64)
            assert O.ola_var8(ptr(a, o), 61) == R.x265ref_var8(ptr(a, o), 61)


@pytest.mark.parametrize("depth", DEPTHS)
def test_mvcost_table_and_exp2fix8(depth):
    O, R = po.oracle(depth), po.ref(depth)
    t1 = np.zeros(4 * 32768 + 1, np.uint16)
    t2 = np.zeros(4 * 32768 + 1, np.uint16)
    O.ola_mvcost_table(ptr(t1))
    R.x265ref_mvcost_table(ptr(t2))
    assert np.array_equal(t1, t2)
    assert O.ola_lambda_int() == R.x265ref_lambda_int()
    for x in np.linspace(-60, 60, 4001):
        assert O.ola_exp2fix8(float(x)) == R.x265ref_exp2fix8(float(x))


@pytest.mark.parametrize("depth", DEPTHS)
def test_propagate_cost(depth):
    """estimateCUPropagateCost (pixel.cpp:848-874; harness recipe pixelharness.cpp check_propagateCost-like: random
    uint16 propagate-in / inter costs, int32 intra costs and inverse qscales, a double fps factor), all lengths incl.
    the vector loop's tail, extremes included"""
    O, R = po.oracle(depth), po.ref(depth)
    rng = np.random.default_rng(5)
    for it in range(60):
        n = int(rng.integers(1, 300))
        pin = rng.integers(0, 65536, n).astype(np.uint16)
        intra = rng.integers(1, 1 << (14 + 2 * (depth > 8)), n).astype(np.int32)
        inter = rng.integers(0, 65536, n).astype(np.uint16)
        invq = rng.integers(1, 4096, n).astype(np.int32)
        if it % 5 == 0:
            intra[: n // 2] = 1
            invq[n // 2:] = 1 << 20     # the int32 product wraps
        fps = C.c_double(float(rng.uniform(0.01, 100.0)) if it % 3 else 1.0)
        a, b = np.zeros(n, np.int32), np.zeros(n, np.int32)
        O.ola_propagate_cost(a.ctypes.data, pin.ctypes.data, intra.ctypes.data, inter.ctypes.data, invq.ctypes.data, C.byref(fps), n)
        R.x265ref_propagate_cost(b.ctypes.data, pin.ctypes.data, intra.ctypes.data, inter.ctypes.data, invq.ctypes.data, C.byref(fps), n)
        assert np.array_equal(a, b), (it, n)


PU_SHAPES = [(4, 4), (8, 8), (16, 16), (32, 32), (64, 64), (4, 8), (8, 4), (16, 8), (8, 16), (16, 12), (12, 16), (16, 4), (4, 16),
             (32, 16), (16, 32), (32, 24), (24, 32), (32, 8), (8, 32), (64, 32), (32, 64), (64, 48), (48, 64), (64, 16), (16, 64)]


@pytest.mark.parametrize("depth", DEPTHS)
def test_pu_sad_satd_all_shapes(depth):
    """the 25 luma PU shapes of pu[].sad / pu[].satd (pixel.cpp:954-1003), check_pixelcmp recipe (pixelharness.cpp:80-99):
    random / all-min / all-max buffers, stride 64 against an odd stride, stepping offsets (oracle side)"""
    O, R = po.oracle(depth), po.ref(depth)
    B = bufs(depth, 23, n=64 * 160)
    for (w, h) in PU_SHAPES:
        for ia, a in enumerate(B[:3]):
            for b in (B[0], B[3 + ia % 2]):
                for it in range(12):
                    oa, ob = (it * 32) % 1500, (it * 37 + 5) % 1500
                    for name in ("sad", "satd"):
                        got = getattr(O, "ola_pu_" + name)(w, h, ptr(a, oa), 64, ptr(b, ob), 67)
                        want = getattr(R, "x265ref_pu_" + name)(w, h, ptr(a, oa), 64, ptr(b, ob), 67)
                        assert got == want, (depth, w, h, name, it)


# ---- explicit weighted-prediction analysis, pixel loops (encoder/weightPrediction.cpp:59-220) ----
def _wp_libs(depth):
    O, R = po.oracle(depth), po.ref(depth)
    V, I, S = C.c_void_p, C.c_int, C.c_ssize_t
    for f in (O.ola_wp_mc_luma, R.x265ref_wp_mc_luma):
        f.argtypes = [V, S, I, I, V, V]
    for f in (O.ola_wp_mc_chroma, R.x265ref_wp_mc_chroma):
        f.argtypes = [V, S, V, I, I, I, I, V]
    for f in (O.ola_wp_cost, R.x265ref_wp_cost):
        f.argtypes = [V, V, V, S, I, I, V, I, I, I, I]
        f.restype = C.c_uint32
    return O, R


def _wp_planes(depth, seed, w, h, mx, my):
    """four border-extended lowres-like planes (smooth + noise) and their origin pointers"""
    dt = po.pixel_dtype(depth)
    rng = np.random.default_rng(seed)
    top = (1 << depth) - 1
    stride = w + 2 * mx
    planes = []
    for k in range(4):
        yy, xx = np.mgrid[0:h, 0:w]
        base = ((xx * 3 + yy * 5 + 17 * k) % (top + 1)).astype(np.int64) // 2 + rng.integers(0, top // 2 + 1, (h, w))
        core = np.clip(base, 0, top).astype(dt)
        planes.append(np.ascontiguousarray(np.pad(core, ((my, my), (mx, mx)), mode="edge")))
    return planes, stride


@pytest.mark.parametrize("depth", DEPTHS)
def test_wp_mc_luma(depth):
    """mcLuma: oracle == the reference's own function on random vectors (in range, far outside, odd and even)"""
    O, R = _wp_libs(depth)
    w, h, mx, my = 88, 56, 48, 40
    planes, stride = _wp_planes(depth, 5, w, h, mx, my)
    org = (C.c_void_p * 4)(*[p.ctypes.data + (my * stride + mx) * p.itemsize for p in planes])
    rng = np.random.default_rng(6)
    n = (w // 8) * (h // 8)
    for spread in (6, 40, 400):
        mvs = rng.integers(-spread, spread + 1, (n, 2)).astype(np.int16)
        a = np.zeros((h, stride), planes[0].dtype)
        b = np.zeros_like(a)
        O.ola_wp_mc_luma(org, stride, w, h, mvs.ctypes.data, a.ctypes.data)
        R.x265ref_wp_mc_luma(org, stride, w, h, mvs.ctypes.data, b.ctypes.data)
        assert np.array_equal(a[:, :w], b[:, :w])


@pytest.mark.parametrize("depth", DEPTHS)
def test_wp_mc_chroma(depth):
    """mcChroma (4:2:0): all four interpolation cases, vectors clipped at the picture edges, the literal availability test"""
    O, R = _wp_libs(depth)
    w, h, mx, my = 96, 64, 40, 40           # chroma analysis area; lowres CU grid 12 x 8 as for a 192 x 128 picture
    planes, stride = _wp_planes(depth, 7, w, h, mx, my)
    src = planes[0]
    p0 = src.ctypes.data + (my * stride + mx) * src.itemsize
    rng = np.random.default_rng(8)
    for wcu, hcu, spread in ((12, 8, 9), (12, 8, 300), (96, 64, 20), (40, 30, 64)):
        mvs = rng.integers(-spread, spread + 1, (max(wcu * h, 4096), 2)).astype(np.int16)
        mvs[::5] &= ~7                       # plenty of full-pel and single-axis vectors
        mvs[1::7, 0] &= ~7
        mvs[2::7, 1] &= ~7
        a = np.zeros((h + 2 * my, stride), src.dtype)
        b = np.zeros_like(a)
        oa = a.ctypes.data + (my * stride + mx) * a.itemsize
        ob = b.ctypes.data + (my * stride + mx) * b.itemsize
        O.ola_wp_mc_chroma(p0, stride, mvs.ctypes.data, wcu, hcu, h, w, oa)
        R.x265ref_wp_mc_chroma(p0, stride, mvs.ctypes.data, wcu, hcu, h, w, ob)
        assert np.array_equal(a, b)
        assert a.any()


@pytest.mark.parametrize("depth", DEPTHS)
def test_wp_cost(depth):
    """weightCost: plain and weighted (the whole range of scales, denominators, offsets), luma (intra-limited) and chroma"""
    O, R = _wp_libs(depth)
    w, h, mx, my = 96, 64, 16, 8
    planes, stride = _wp_planes(depth, 9, w, h, mx, my)
    f, r = planes[0], planes[1]
    pf = f.ctypes.data + (my * stride + mx) * f.itemsize
    pr = r.ctypes.data + (my * stride + mx) * r.itemsize
    tmp = np.zeros((h + 1, stride), f.dtype)
    rng = np.random.default_rng(10)
    intra = rng.integers(0, 1200 << (depth - 8), (w // 8) * (h // 8)).astype(np.int32)
    cases = [(0, 0, 0, 0)] + [(1, int(s), int(d), int(o)) for s, d, o in zip(rng.integers(0, 128, 40), rng.integers(0, 8, 40), rng.integers(-128, 128, 40))]
    for luma in (True, False):
        ic = intra.ctypes.data if luma else None
        for wt, s, d, o in cases:
            a = O.ola_wp_cost(pf, pr, tmp.ctypes.data, stride, w, h, ic, wt, s, d, o)
            b = R.x265ref_wp_cost(pf, pr, tmp.ctypes.data, stride, w, h, ic, wt, s, d, o)
            assert a == b, (luma, wt, s, d, o)


# ---- full-resolution PU motion search (encoder/motion.cpp:571-1172, every pattern and sub-pel level) ----
from harness import me_cases as mc


def _me_libs(depth):
    O, R = po.oracle(depth), po.ref(depth)
    V, I, S = C.c_void_p, C.c_int, C.c_ssize_t
    O.ola_motion_estimate_batch.argtypes = [I, I, I, I, V, S, V, S, V, I, V, V]
    O.ola_motion_estimate_batch.restype = None
    R.x265ref_motion_estimate_batch.argtypes = [I, I, I, I, I, V, S, V, S, I, V, V, V]
    R.x265ref_motion_estimate_batch.restype = None
    R.x265ref_mvcost_table_qp.argtypes = [I, V]
    return O, R


def _me_compare(depth, method, subme, w, h, n, seed, qp=30, W=160, H=128, **kw):
    O, R = _me_libs(depth)
    fenc, ref, stride, _ = mc.planes(depth, W, H, seed)
    its = mc.items(W, H, w, h, n, seed + 1, **kw)
    lut = np.zeros(131073, np.uint16)
    R.x265ref_mvcost_table_qp(qp, ptr(lut))
    want = (mc.MeResult * n)(); got = (mc.MeResult * n)()
    R.x265ref_motion_estimate_batch(method, subme, qp, w, h, ptr(fenc), stride, ptr(ref), stride, n, C.addressof(its), C.addressof(want), None)
    O.ola_motion_estimate_batch(method, subme, w, h, ptr(fenc), stride, ptr(ref), stride, ptr(lut, 65536), n, C.addressof(its), C.addressof(got))
    a, b = mc.results_list(want, n), mc.results_list(got, n)
    bad = [i for i in range(n) if a[i] != b[i]]
    assert not bad, (depth, method, subme, w, h, bad[:5], [(a[i], b[i]) for i in bad[:3]])
    return a


@pytest.mark.parametrize("depth", DEPTHS)
@pytest.mark.parametrize("method", ["dia", "hex", "umh", "star", "full"])
def test_motion_estimate_patterns(depth, method):
    """every integer pattern x every sub-pel level on 16x16 and 8x8 PUs, windows that cut the patterns included"""
    m = mc.METHODS[method]
    moved = 0
    for subme in range(8):
        for (w, h) in ((16, 16), (8, 8)):
            n = 12 if method == "full" else 60
            kw = dict(merange_choices=(4, 9)) if method == "full" else {}
            res = _me_compare(depth, m, subme, w, h, n, 100 + subme * 7 + w, **kw)
            moved += sum(1 for r in res if (r[0], r[1]) != (0, 0))
    assert moved > 50


@pytest.mark.parametrize("depth", DEPTHS)
def test_motion_estimate_all_shapes(depth):
    """the 25 luma PU shapes (4x4 is not a legal inter PU: skipped, motion.cpp:168) with STAR and UMH at the default sub-pel level 2 and at 5"""
    for (w, h) in mc.PU_SHAPES:
        if (w, h) == (4, 4):
            continue
        for method in ("star", "umh", "hex"):
            for subme in (2, 5):
                _me_compare(depth, mc.METHODS[method], subme, w, h, 16, 300 + w * 3 + h, W=192, H=160)


@pytest.mark.parametrize("depth", DEPTHS)
def test_motion_estimate_far_motion(depth):
    """a reference moved by (37, -22) with predictors at zero: long walks, the 16-point grid of UMH, STAR's raster refinement"""
    O, R = _me_libs(depth)
    for method in ("star", "umh", "hex", "dia"):
        for (w, h) in ((16, 16), (32, 32)):
            fenc, ref, stride, _ = mc.planes(depth, 256, 192, 77, motion=(31, -22), noise=2)
            its = mc.items(256, 192, w, h, 40, 78, merange_choices=(32, 57), tight=0.1)
            for i in range(40):
                if i % 2 == 0:
                    its[i].qmvp[0] = its[i].qmvp[1] = 0
                    its[i].mvmin[0], its[i].mvmin[1], its[i].mvmax[0], its[i].mvmax[1] = -24, -24, 24, 24
                    its[i].offset = mc.MARGIN * stride + mc.MARGIN + (64 + (i % 5) * 8) * stride + 96 + (i % 7) * 8
                    its[i].mvmin[0] = -57; its[i].mvmin[1] = -57; its[i].mvmax[0] = 57; its[i].mvmax[1] = 57
            lut = np.zeros(131073, np.uint16)
            R.x265ref_mvcost_table_qp(26, ptr(lut))
            want = (mc.MeResult * 40)(); got = (mc.MeResult * 40)()
            R.x265ref_motion_estimate_batch(mc.METHODS[method], 3, 26, w, h, ptr(fenc), stride, ptr(ref), stride, 40, C.addressof(its), C.addressof(want), None)
            O.ola_motion_estimate_batch(mc.METHODS[method], 3, w, h, ptr(fenc), stride, ptr(ref), stride, ptr(lut, 65536), 40, C.addressof(its), C.addressof(got))
            assert mc.results_list(want, 40) == mc.results_list(got, 40), (method, w, h)
            if method in ("star", "umh"):
                far = sum(1 for r in mc.results_list(want, 40) if abs(r[0]) > 80)
                assert far >= 10, (method, far)
