"""Explicit weighted-prediction analysis on the GPU: x265cu_wp_prepare / x265cu_wp_cost against the oracle's
restatement of mcLuma / mcChroma / weightCost (encoder/weightPrediction.cpp:59-220), which tests/test_oracle_vs_ref.py pins
against the reference's own compiled functions.  The planes the kernels read are the ones the pre-lookahead left on the
device (lowres planes of the slots, compact chroma copies); the oracle reads the host copies of the same pictures."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mods(built):
    from oracle import pyoracle as po
    from src.x265_b200 import abi
    if abi.lib_cu().x265cu_device_count() < 1:
        pytest.fail("no CUDA device visible: the GPU tests must run on the B200 box (there is no CPU fallback)")
    return po, abi


def _chroma_variant(a, t, depth):
    """the synthetic chroma is static; give every frame its own so that the motion compensation has something to move"""
    rng = np.random.default_rng(1000 + t)
    top = (1 << depth) - 1
    b = a.astype(np.int64) + rng.integers(-40 << (depth - 8), (40 << (depth - 8)) + 1, a.shape)
    return np.clip(b, 0, top).astype(a.dtype)


@pytest.mark.parametrize("depth,w,h", [(8, 320, 192), (8, 360, 208), (10, 320, 192), (8, 1280, 720)])
def test_wp_analysis_kernels(mods, depth, w, h):
    po, abi = mods
    O = po.oracle(depth)
    V, I, S = C.c_void_p, C.c_int, C.c_ssize_t
    O.ola_wp_mc_luma.argtypes = [V, S, I, I, V, V]
    O.ola_wp_mc_chroma.argtypes = [V, S, V, I, I, I, I, V]
    O.ola_wp_cost.argtypes = [V, V, V, S, I, I, V, I, I, I, I]
    O.ola_wp_cost.restype = C.c_uint32
    L = abi.lib_cu()
    L.x265cu_wp_prepare.argtypes = [V, I, I, I, V, V]
    L.x265cu_wp_cost.argtypes = [V, I, V, V]
    dt = po.pixel_dtype(depth)
    la = abi.Lookahead(w, h, depth, 3, 10, 0, 16, 1, 1, 1.0, 0, 0, 8)
    frames, pics = [], []
    for t in (0, 3):
        y, u, v = po.synth_padded(O, depth, w, h, t, 12, 77)
        ub, vb = u.base if u.base is not None else u, v.base if v.base is not None else v
        u[:] = _chroma_variant(u, t, depth); v[:] = _chroma_variant(v, t + 50, depth)
        f = la.frame_alloc()
        la.pre_lookahead(f, y, u, v, t, planes_back=True)
        frames.append(f); pics.append((y, u, v, ub, vb))
    la.sync()
    g = la.info() if hasattr(la, "info") else None
    wcu, hcu, ncu = la.wCU, la.hCU, la.nCU
    stride, plane_size, pad_off = la.stride, la.planeSize, la.padOffset
    lw, ll = wcu * 8, hcu * 8
    slots = [abi.lib_host().x265cuh_frame_slot(f) for f in frames]
    planes = [la.array(f, 0, dtype=dt).reshape(4, plane_size) for f in frames]
    rng = np.random.default_rng(depth * 1000 + w)
    intra = rng.integers(0, 2000 << (depth - 8), ncu).astype(np.int32)
    cands = [(0, 0, 0, 0), (1, 64, 6, 0), (1, 70, 6, -3), (1, 30, 5, 12), (1, 127, 7, -128), (1, 1, 0, 127), (1, 90, 7, 5)]
    items = (abi.WeightItem * len(cands))()
    for i, (wt, s, d, o) in enumerate(cands):
        items[i].fenc = 0; items[i].ref = 0; items[i].weighted = wt; items[i].scale = s; items[i].denom = d; items[i].offset = o
    costs = (C.c_uint32 * len(cands))()
    tmp = np.zeros(max(plane_size, (h // 2 + 200) * (w // 2 + 200)) + 4096, dt)

    def check(plane, mvs):
        mp = mvs.ctypes.data if mvs is not None else None
        assert L.x265cu_wp_prepare(la.ctx, slots[1], slots[0], plane, mp, intra.ctypes.data if plane == 0 else None) == 0, L.x265cu_last_error(la.ctx)
        assert L.x265cu_wp_cost(la.ctx, len(cands), items, costs) == 0, L.x265cu_last_error(la.ctx)
        if plane == 0:
            fp = planes[1][0].ctypes.data + pad_off * planes[1].itemsize
            rorg = (C.c_void_p * 4)(*[planes[0][k].ctypes.data + pad_off * planes[0].itemsize for k in range(4)])
            rp = planes[0][0].ctypes.data + pad_off * planes[0].itemsize
            mc = np.zeros(plane_size, dt)
            if mvs is not None:
                O.ola_wp_mc_luma(rorg, stride, lw, ll, mp, mc.ctypes.data)
                rp = mc.ctypes.data
            want = [O.ola_wp_cost(fp, rp, tmp.ctypes.data, stride, lw, ll, intra.ctypes.data, *c) for c in cands]
        else:
            cw, ch = ((w >> 4) << 4) >> 1, ((h >> 4) << 4) >> 1
            m = 48
            src = pics[0][plane][: h // 2, : w // 2]
            fen = pics[1][plane][: h // 2, : w // 2]
            ext = np.ascontiguousarray(np.pad(src, m, mode="edge"))          # extendPicBorder
            fex = np.ascontiguousarray(np.pad(fen, m, mode="edge"))
            es = ext.shape[1]
            org = ext.ctypes.data + (m * es + m) * ext.itemsize
            forg = fex.ctypes.data + (m * es + m) * fex.itemsize
            rp = org
            if mvs is not None:
                mc = np.zeros_like(ext)
                O.ola_wp_mc_chroma(org, es, mp, lw >> 3, ll >> 3, ch, cw, mc.ctypes.data + (m * es + m) * mc.itemsize)
                rp = mc.ctypes.data + (m * es + m) * mc.itemsize
            want = [O.ola_wp_cost(forg, rp, tmp.ctypes.data, es, cw, ch, None, *c) for c in cands]
        assert list(costs) == want, (plane, mvs is not None)

    for plane in (0, 1, 2):
        check(plane, None)
        for spread in (5, 60, 3000):
            mvs = rng.integers(-spread, spread + 1, (max(ncu, (lw >> 3) * (h // 2)), 2)).astype(np.int16)
            mvs[::3] &= ~7
            mvs[1::5, 0] &= ~7
            mvs[2::5, 1] &= ~7
            check(plane, mvs)
    for f in frames:
        la.frame_free(f)
    la.close()
