"""GPU path against the ORACLE on estimates the reference clips never ask for: the largest table sizes (bframes = 16:
distances up to 17 in both lists), every kind of estimate in one batch, repeated requests (cached), cuTree steps over
those distances, and the degenerate calls (empty lists).  The golden traces pin the oracle and the GPU on what x265
really runs; this pins the GPU on the rest of the domain the ABI accepts, with the oracle (itself pinned against the
reference on every trace) as the checker."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mods(built):
    from harness import replay
    from oracle import pyoracle as po
    from src.x265_b200 import abi
    if abi.lib_cu().x265cu_device_count() < 1:
        pytest.fail("no CUDA device visible: the GPU tests must run on the B200 box (there is no CPU fallback)")
    return replay, po, abi


class FakeTrace:
    def __init__(self, cfg, nframes):
        self.cfg = cfg
        self.events = [("P", dict(poc=t)) for t in range(nframes)]

    def jobs(self):
        return iter(())


def _cfg(depth, w, h, nframes, bframes, weightp=1):
    return dict(width=w, height=h, depth=depth, nframes=nframes, seed=4242, bframes=bframes, lookahead=nframes, badapt=2, weightp=weightp,
                aqmode=1, cutree=1, numCoopSlices=1, numRowsPerSlice=(h // 2 + 7) // 8, bFrameBias=0, pool=16, aqStrength=1.0,
                fpsNum=30, fpsDenom=1, qCompress=0.6, weightedBiPred=0, vbvBufferSize=0)


def _oracle_frames(po, cfg):
    lib = po.oracle(cfg["depth"])
    frames = {}
    for poc in range(cfg["nframes"]):
        f = po.OFrame(lib, cfg["depth"], cfg["width"], cfg["height"], 96, 80, cfg["bframes"], True)
        y, u, v = po.synth_padded(lib, cfg["depth"], cfg["width"], cfg["height"], poc, cfg["nframes"], cfg["seed"])
        lib.ola_frame_init(f.p, y.ctypes.data, y.strides[0] // y.itemsize, poc)
        lib.ola_aq_frame(f.p, y.ctypes.data, y.strides[0] // y.itemsize, u.ctypes.data, v.ctypes.data, u.strides[0] // u.itemsize,
                         cfg["aqmode"], cfg["aqStrength"], cfg["weightp"])
        lib.ola_intra_estimate(f.p, 1 if cfg["depth"] == 8 else 16)
        frames[poc] = f
    return lib, frames


def _compare_job(po, la, gf, of, d0, d1, tag):
    s = la.scalars(gf, d0, d1)
    assert s["costEst"] == int(of.c.costEst[d0][d1]), tag + " costEst"
    assert s["costEstAq"] == int(of.c.costEstAq[d0][d1]), tag + " costEstAq"
    if d1 == 0:
        assert s["intraMbs"] == int(of.c.intraMbs[d0]), tag + " intraMbs"
    assert np.array_equal(la.array(gf, 4, d0, d1, np.uint16), of.lowres_costs(d0, d1)), tag + " lowresCosts"
    assert np.array_equal(la.array(gf, 5, d0, d1, np.int32), of.row_satds(d0, d1)), tag + " rowSatds"
    for lst, d in ((0, d0), (1, d1)):
        if d > 0:
            assert np.array_equal(la.array(gf, 6, lst, d, np.int16), of.mvs(lst, d)), tag + " mvs%d" % lst
            assert np.array_equal(la.array(gf, 7, lst, d, np.int32), of.mv_costs(lst, d)), tag + " mvCosts%d" % lst


@pytest.mark.parametrize("depth", [8, 10])
def test_max_distances_bframes16(mods, depth):
    replay, po, abi = mods
    N, BF = 19, 16
    cfg = _cfg(depth, 416, 240, N, BF)
    lib, of = _oracle_frames(po, cfg)
    octx = lib.ola_ctx_create(cfg["bFrameBias"], 1, cfg["numRowsPerSlice"])
    r = replay.CuReplay(FakeTrace(cfg, N), check=False)
    try:
        r.run()                                      # pre-lookahead of all frames (one list)
        la = r.la
        for t in range(N):
            assert la.crc(r.frames[t], 1) == po.crc(of[t].intra_cost()), "intraCost %d" % t
            assert la.crc(r.frames[t], 3) == po.crc(of[t].inv_qscale()), "invQscale %d" % t
        # single estimates, in this order on both sides (the 0x7FFF sentinels decide what each one searches)
        singles = [(0, 17, 17), (0, 17, 8), (0, 17, 1), (0, 17, 16), (1, 18, 18), (1, 18, 2), (2, 3, 3), (0, 17, 8), (16, 18, 17), (0, 9, 8)]
        frames = [r.frames[t] for t in range(N)]
        for (p0, p1, b) in singles:
            d0, d1 = b - p0, p1 - b
            used = po.Weight()
            want = lib.ola_estimate(octx, of[b].p, of[p0].p, of[p1].p, d0, d1, -1, -1, 1, cfg["weightp"], None, C.byref(used))
            got = la.estimate(frames, [(p0, p1, b)], False)[0]
            assert got == int(want), (p0, b, p1)
            _compare_job(po, la, r.frames[b], of[b], d0, d1, "single %d/%d/%d" % (p0, b, p1))
        # one batch with every kind of estimate: P and B, all distances from one frame, independent as the reference's are
        batch = [(b - d, b, b) for b in range(4, N) for d in (1, 2, 3)] + [(3, 9, b) for b in range(4, 9)] + [(0, 18, 9), (2, 18, 10)]
        seen, uniq = set(), []
        for p0, p1, b in batch:     # a batch may search a field only once (the reference's batches are independent)
            keys = {(b, 0, b - p0)} | ({(b, 1, p1 - b)} if p1 > b else set())
            if not (keys & seen):
                seen |= keys
                uniq.append((p0, p1, b))
        for (p0, p1, b) in uniq:
            used = po.Weight()
            lib.ola_estimate(octx, of[b].p, of[p0].p, of[p1].p, b - p0, p1 - b, -1, -1, 0, cfg["weightp"], None, C.byref(used))
        la.estimate(frames, uniq, True)
        for (p0, p1, b) in uniq:
            _compare_job(po, la, r.frames[b], of[b], b - p0, p1 - b, "batch %d/%d/%d" % (p0, b, p1))
        # cuTree over the long distances
        for t in (17, 0, 8):
            la.cutree_zero(r.frames[t])
            lib.ola_cutree_zero(of[t].p)
        for (p0, p1, b, ref) in [(0, 17, 17, 1), (0, 17, 8, 1), (0, 17, 1, 0), (0, 17, 16, 0), (0, 9, 8, 0)]:
            la.cutree_propagate(frames, p0, p1, b, ref, 1.0 / 30)
            lib.ola_estimate_cu_propagate(of[b].p, of[p0].p, of[p1].p, b - p0, p1 - b, ref, 1.0 / 30, 30, 1, 0)
        la.cutree_finish(r.frames[0], 1.0 / 30, 0)
        lib.ola_cutree_finish(of[0].p, 1.0 / 30, 30, 1, 0, 5.0 * (1.0 - 0.6))
        for t in (0, 8, 17, 9):
            assert np.array_equal(la.array(r.frames[t], 8, 0, 0, np.uint16), of[t].propagate_cost()), "propagateCost %d" % t
        assert np.array_equal(la.array(r.frames[0], 9, 0, 0, np.float64), of[0].qp_cutree_offset()), "qpCuTreeOffset"
    finally:
        r.close()
        lib.ola_ctx_destroy(octx)
        for f in of.values():
            f.close()


def test_empty_calls(mods):
    """n = 0 everywhere: nothing happens, nothing fails"""
    replay, po, abi = mods
    L = abi.lib_cu()
    la = abi.Lookahead(320, 192, 8, 3, 10, 0, 0, 1, 1, 1.0)
    try:
        assert L.x265cu_estimate_batch(la.ctx, 0, None, None) == 0
        assert L.x265cu_weight_cost_batch(la.ctx, 0, None, None) == 0
        assert L.x265cu_cutree_run(la.ctx, 0, None, 0, None, None) == 0
        L.x265cu_frame_init_var_batch.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.x265cu_intra_batch.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.x265cu_pre_lookahead_batch.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        assert L.x265cu_frame_init_var_batch(la.ctx, 0, None) == 0
        assert L.x265cu_intra_batch(la.ctx, 0, None, None) == 0
        assert L.x265cu_pre_lookahead_batch(la.ctx, 0, None, None, None, None) == 0
        assert L.x265cu_pixelcmp_batch(la.ctx, 1, la.ctx, 0, 64, la.ctx, 0, 64, 0, None, None, None) == 0
        assert la.estimate([], [], True) == []
        assert L.x265cu_sync(la.ctx) == 0
    finally:
        la.close()
