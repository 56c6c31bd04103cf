"""Deterministic inputs for the full-resolution PU motion search (encoder/motion.cpp:571-1172):
a source plane and a reference plane with margins (integer arithmetic only, so every machine builds the same bytes) and
lists of search items in the layout of include/x265cu.h's x265cu_me_item.

TEST / BENCH INFRASTRUCTURE: used by tests/test_oracle_vs_ref.py (oracle == reference), tests/test_gpu_me.py (CUDA ==
oracle == committed reference results), tools/me_bench.py and oracle/gen_me_golden.py."""
import ctypes as C

import numpy as np

PU_SHAPES = [(4, 4), (8, 8), (16, 16), (32, 32), (64, 64), (8, 4), (4, 8), (16, 8), (8, 16), (32, 16), (16, 32), (64, 32), (32, 64),
             (16, 12), (12, 16), (16, 4), (4, 16), (32, 24), (24, 32), (32, 8), (8, 32), (64, 48), (48, 64), (64, 16), (16, 64)]
METHODS = {"dia": 0, "hex": 1, "umh": 2, "star": 3, "full": 4}
MARGIN = 96          # samples around the visible picture on every side (x265: CTU + 32 / CTU + 16)
SLACK = 24           # how far a search window may reach outside the picture (the patterns overshoot by <= 6, sub-pel by 5)


class MeItem(C.Structure):
    _fields_ = [("offset", C.c_int64), ("mvmin", C.c_int16 * 2), ("mvmax", C.c_int16 * 2), ("qmvp", C.c_int16 * 2),
                ("numCandidates", C.c_int16), ("merange", C.c_int16), ("mvc", (C.c_int16 * 2) * 12)]


class MeResult(C.Structure):
    _fields_ = [("mv", C.c_int16 * 2), ("cost", C.c_int32)]


def _box(a, r):
    """integer box blur of radius r (wraps around: only used on the unpadded texture)"""
    acc = np.zeros(a.shape, np.int64)
    for dy in range(-r, r + 1):
        for dx in range(-r, r + 1):
            acc += np.roll(np.roll(a, dy, 0), dx, 1)
    return acc // ((2 * r + 1) ** 2)


def planes(depth, width, height, seed, motion=(5, -3), noise=6):
    """(fenc, ref, stride, origin): two padded planes (edge-replicated margins).  ref is the source texture moved by `motion`
    full-pel samples plus a half-sample blend and noise, so that integer patterns travel and sub-pel steps matter."""
    rng = np.random.default_rng(seed)
    mx = (1 << depth) - 1
    big = rng.integers(0, 256, (height + 64, width + 64)).astype(np.int64)
    tex = (_box(big, 2) * 3 + _box(big, 6) * 2 + big) // 6
    tex = (tex - tex.min()) * mx // max(1, int(tex.max() - tex.min()))
    src = tex[32:32 + height, 32:32 + width]
    dx, dy = motion
    a = tex[32 + dy:32 + dy + height, 32 + dx:32 + dx + width]
    b = tex[32 + dy:32 + dy + height, 33 + dx:33 + dx + width]
    moved = (3 * a + b + 2) // 4 + rng.integers(-noise, noise + 1, (height, width)) * (1 << (depth - 8))
    moved = np.clip(moved, 0, mx)
    dt = np.uint8 if depth == 8 else np.uint16
    out = []
    for p in (src, moved):
        out.append(np.ascontiguousarray(np.pad(p.astype(dt), MARGIN, mode="edge")))
    stride = width + 2 * MARGIN
    return out[0], out[1], stride, MARGIN * stride + MARGIN


def items(width, height, w, h, n, seed, merange_choices=(8, 16, 32, 57), max_cand=5, tight=0.25, stride=None):
    """n search items for PUs of w x h inside a width x height picture"""
    rng = np.random.default_rng(seed)
    stride = stride or (width + 2 * MARGIN)
    arr = (MeItem * n)()
    for i in range(n):
        x = int(rng.integers(0, (width - w) // 4 + 1)) * 4
        y = int(rng.integers(0, (height - h) // 4 + 1)) * 4
        it = arr[i]
        it.offset = MARGIN * stride + MARGIN + y * stride + x
        mer = int(rng.choice(merange_choices))
        it.merange = mer
        spread = int(rng.choice([0, 6, 40, 120]))
        qx, qy = int(rng.integers(-spread, spread + 1)), int(rng.integers(-spread, spread + 1))
        if rng.random() < 0.5:
            qx += 20; qy -= 12                                   # near the true motion (5, -3)
        it.qmvp[0], it.qmvp[1] = qx, qy
        # setSearchRange-like window: the predictor +- merange, limited to the picture + SLACK
        lox, hix = -x - SLACK, width - x - w + SLACK
        loy, hiy = -y - SLACK, height - y - h + SLACK
        cx, cy = qx >> 2, qy >> 2
        if rng.random() < tight:
            # windows that cut the patterns: a few samples on one or more sides
            mnx, mxx = cx - int(rng.integers(0, 6)), cx + int(rng.integers(0, mer + 1))
            mny, mxy = cy - int(rng.integers(0, mer + 1)), cy + int(rng.integers(0, 6))
        else:
            mnx, mxx, mny, mxy = cx - mer, cx + mer, cy - mer, cy + mer
        mnx, mxx = max(mnx, lox), min(mxx, hix)
        mny, mxy = max(mny, loy), min(mxy, hiy)
        if mnx > mxx:
            mnx = mxx = max(lox, min(hix, cx))
        if mny > mxy:
            mny = mxy = max(loy, min(hiy, cy))
        it.mvmin[0], it.mvmin[1], it.mvmax[0], it.mvmax[1] = mnx, mny, mxx, mxy
        nc = int(rng.integers(0, max_cand + 1))
        it.numCandidates = nc
        for k in range(nc):
            kind = rng.random()
            if kind < 0.3:
                it.mvc[k][0], it.mvc[k][1] = qx, qy
            elif kind < 0.4:
                it.mvc[k][0], it.mvc[k][1] = 0, 0
            else:
                it.mvc[k][0] = qx + int(rng.integers(-30, 31))
                it.mvc[k][1] = qy + int(rng.integers(-30, 31))
    return arr


def results_list(res, n):
    return [(int(res[i].mv[0]), int(res[i].mv[1]), int(res[i].cost)) for i in range(n)]


def make_lut(lam_q2=20):
    """a BitCost-shaped mvcost table every machine rebuilds bit for bit: lambda * (exp-Golomb length of |d|) in pure integers,
    saturated like bitcost.cpp:57 (the real tables are float-made on the host and are an INPUT of the search)"""
    bits = np.array([2 * (int(v + 1).bit_length() - 1) + 1 for v in range(65537)], dtype=np.int64)
    half = np.minimum((bits * lam_q2) >> 2, 32767).astype(np.uint16)
    lut = np.zeros(131073, np.uint16)
    lut[65536:] = half
    lut[:65537] = half[::-1]
    return lut


# (method, subme, w, h, n, seed, picture W, H, item kwargs): the committed reference results (tests/golden/me_golden.json,
# made by oracle/gen_me_golden.py from the unmodified reference) cover every pattern, every sub-pel level, every PU shape
def golden_cases():
    cases = []
    seed = 1000
    for method in ("dia", "hex", "umh", "star"):
        for subme in range(8):
            for (w, h) in ((16, 16), (8, 8)):
                seed += 1
                cases.append(dict(method=method, subme=subme, w=w, h=h, n=24, seed=seed, W=160, H=128, kw={}))
    for subme in (1, 5):
        seed += 1
        cases.append(dict(method="full", subme=subme, w=16, h=16, n=8, seed=seed, W=160, H=128, kw=dict(merange_choices=(4, 9))))
    for (w, h) in PU_SHAPES:
        if (w, h) == (4, 4):
            continue
        for method in ("star", "umh"):
            seed += 1
            cases.append(dict(method=method, subme=2 if method == "star" else 5, w=w, h=h, n=10, seed=seed, W=192, H=160, kw={}))
    # far motion: long walks, UMH's grid, STAR's raster refinement
    for method in ("star", "umh", "hex"):
        seed += 1
        cases.append(dict(method=method, subme=3, w=16, h=16, n=24, seed=seed, W=256, H=192, kw=dict(merange_choices=(32, 57), tight=0.1),
                          motion=(31, -22), noise=2, zero_pred=True))
    return cases


def build_case(case, depth):
    """(fenc, ref, stride, items) of a golden case"""
    fenc, ref, stride, _ = planes(depth, case["W"], case["H"], case["seed"], motion=tuple(case.get("motion", (5, -3))), noise=case.get("noise", 6))
    its = items(case["W"], case["H"], case["w"], case["h"], case["n"], case["seed"] + 1, **case["kw"])
    if case.get("zero_pred"):
        for i in range(0, case["n"], 2):
            its[i].qmvp[0] = its[i].qmvp[1] = 0
            its[i].offset = MARGIN * stride + MARGIN + (64 + (i % 5) * 8) * stride + 96 + (i % 7) * 8
            its[i].mvmin[0] = its[i].mvmin[1] = -57
            its[i].mvmax[0] = its[i].mvmax[1] = 57
    return fenc, ref, stride, its


def oracle_me(depth):
    """the CPU oracle's entry for this path (test / bench infrastructure, like this module)"""
    from oracle import pyoracle as po
    O = po.oracle(depth)
    V, I, S = C.c_void_p, C.c_int, C.c_ssize_t
    O.ola_motion_estimate_batch.argtypes = [I, I, I, I, V, S, V, S, V, I, V, V]
    O.ola_motion_estimate_batch.restype = None
    return O


def oracle_run(O, case, depth, lut, built=None):
    fenc, ref, stride, its = built or build_case(case, depth)
    n = case["n"]
    got = (MeResult * n)()
    O.ola_motion_estimate_batch(METHODS[case["method"]], case["subme"], case["w"], case["h"], fenc.ctypes.data, stride, ref.ctypes.data, stride,
                                lut.ctypes.data + 65536 * 2, n, C.addressof(its), C.addressof(got))
    return results_list(got, n)
