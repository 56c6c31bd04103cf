"""GPU parity tests proper: the CUDA path, called through the host layer and the C ABI, must
reproduce bit-for-bit what the UNMODIFIED reference produced (CRC32 of every output array of every
pre-lookahead frame and every frame-cost estimate in the golden traces) and what the oracle
computes for the pixel primitives on the reference TestBench's own input recipes."""
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


def _replay(mods, name, max_events=None):
    replay, po, abi = mods
    res = replay.replay_trace(name, max_events=max_events)
    assert res["jobs"] > 0 and res["frames"] > 0
    assert res["propagates"] > 0      # the cuTree steps of the trace were replayed (and compared) too
    assert not res["mismatches"], "%s: %d mismatches, first: %r" % (name, len(res["mismatches"]), res["mismatches"][:5])
    return res


@pytest.mark.parametrize("name", ["tiny8", "odd8", "tiny10"])
def test_replay_small(mods, name):
    _replay(mods, name)


@pytest.mark.parametrize("name", ["c0_720p", "pool3_720p", "c0_720p10"])
def test_replay_720p(mods, name):
    """720p: cooperative slices (sliced searches), weighted prediction, AQ, 8- and 10-bit"""
    _replay(mods, name)


@pytest.mark.parametrize("mode", ["0", "1"])
@pytest.mark.parametrize("name", ["c0_720p", "c0_720p10", "odd8"])
def test_replay_search_paths(mods, monkeypatch, name, mode):
    """every search through ONE kernel path (the default picks per search): 0 = the plain wavefront kernel only,
    1 = the speculative path only (refine iterations + memoised wide commit, in-batch seed waves)"""
    monkeypatch.setenv("X265CU_SEARCH_MODE", mode)
    _replay(mods, name)


@pytest.mark.parametrize("rows", ["2", "8"])
def test_replay_row_groups(mods, monkeypatch, rows):
    """other row-group sizes of both search kernels (hand-off between CTAs through global memory at other rows)"""
    monkeypatch.setenv("X265CU_PLAIN_ROWS", rows)
    monkeypatch.setenv("X265CU_SEARCH_MODE", "2")
    monkeypatch.setenv("X265CU_SPEC_MAX_PLANS", "64")
    _replay(mods, "c0_720p")


@pytest.mark.parametrize("name,win", [("c0_720p", "0"), ("c0_720p10", "1"), ("odd8", "1")])
def test_replay_plain_window_variants(mods, monkeypatch, name, win):
    """the plain kernel's shared-memory window variant is on for 8-bit and off for 16-bit samples by default:
    cover the other setting of each (and odd picture sizes through the window)"""
    monkeypatch.setenv("X265CU_PLAIN_WIN", win)
    monkeypatch.setenv("X265CU_SEARCH_MODE", "0")
    _replay(mods, name)


@pytest.mark.parametrize("name,env", [("c0_720p", {"X265CU_PLAIN_OCT": "1"}), ("c0_720p10", {"X265CU_PLAIN_OCT": "1"}),
                                      ("pool3_720p", {"X265CU_PLAIN_OCT": "1", "X265CU_OCT_WARPS": "1", "X265CU_OCT_SLACK": "3", "X265CU_OCT_FULL_WARPS": "1"}),
                                      ("odd8", {"X265CU_PLAIN_OCT": "1", "X265CU_PLAIN_WIN": "0"}),
                                      ("c0_720p", {"X265CU_PLAIN_ONESHOT": "1"}), ("c0_720p10", {"X265CU_PLAIN_ONESHOT": "1", "X265CU_PLAIN_WIN": "1"}),
                                      ("odd8", {"X265CU_PLAIN_ONESHOT": "1"}),
                                      ("c0_720p", {"X265CU_PLAIN_TMA": "1"}), ("c0_720p10", {"X265CU_PLAIN_TMA": "1", "X265CU_PLAIN_WIN": "1"}),
                                      ("odd8", {"X265CU_PLAIN_TMA": "1"}), ("pool3_720p", {"X265CU_PLAIN_TMA": "1", "X265CU_PLAIN_ROWS": "8"})])
def test_replay_experimental_search_kernels(mods, monkeypatch, name, env):
    """kernels kept for the record, off by default (profiles/README.md: measured slower): the octet wavefront kernel (an octet
    per CU, four CU rows per warp in lock step; other band counts, with the keep-a-distance wait, without the window) and the
    plain kernel's one-shot burst (la_fast_path) and its TMA-staged, double-buffered windows (cp.async.bulk.tensor.3d +
    mbarrier; weighted references fall back to the load path inside the same launch); all must still be bit-exact"""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    monkeypatch.setenv("X265CU_SEARCH_MODE", "0")
    _replay(mods, name)


def test_replay_without_lookahead_cache(mods, monkeypatch):
    """the host layer's look-ahead estimate cache off: every non-batch estimate computed on request"""
    monkeypatch.setenv("X265CU_LOOKAHEAD_CACHE", "0")
    _replay(mods, "c0_720p")
    _replay(mods, "pool3_720p")


def test_replay_without_rule_based_prediction(mods, monkeypatch):
    """the cache with its history only (no trellis-segment prediction for requests the history cannot foresee)"""
    monkeypatch.setenv("X265CU_TRELLIS_AHEAD", "0")
    _replay(mods, "c0_720p")
    _replay(mods, "pool3_720p")


@pytest.mark.parametrize("var,val", [("X265CU_PRE_PIPELINE", "0"), ("X265CU_MAPPED_RESULTS", "0"), ("X265CU_DEFER_PLANES", "0")])
def test_replay_transfer_variants(mods, monkeypatch, var, val):
    """the other setting of each host<->device transfer strategy: pre-lookahead list as three calls instead of the pipelined
    one, result arrays through pinned staging + host memcpy instead of straight into mapped destinations, plane copy-backs
    issued at once instead of held back until the next estimate batch"""
    monkeypatch.setenv(var, val)
    _replay(mods, "c0_720p")
    _replay(mods, "tiny10")


def test_replay_frame_var_quad_form(mods, monkeypatch):
    """the variance kernel's other wide form (a quad per 16x16 block instead of a pair of lanes), 8 and 10 bit"""
    monkeypatch.setenv("X265CU_VAR_PAIR", "0")
    _replay(mods, "c0_720p")
    _replay(mods, "tiny10")


def test_replay_config1_1080p(mods):
    """BASELINE.json configs[1]: 1080p, b-adapt 2, rc-lookahead 40, cuTree on (the bench workload)"""
    _replay(mods, "c1_1080p")


def test_replay_config0_1080p(mods):
    _replay(mods, "c0_1080p")


def test_replay_config2_4k(mods):
    """configs[2]: 4K, rc-lookahead 40, bframes 8"""
    _replay(mods, "c2_4k")


def test_replay_config3_4k_10bit(mods):
    """configs[3]: 10-bit 4K, --preset slow"""
    _replay(mods, "c3_4k10")


def _pixel_buffers(po, depth, seed):
    """the three TestBench buffers: random, all-min, all-max (test/pixelharness.cpp:30-62)"""
    dt = po.pixel_dtype(depth)
    rng = np.random.default_rng(seed)
    n = 64 * 64 * 4
    mx = (1 << depth) - 1
    return [rng.integers(0, mx + 1, n).astype(dt), np.zeros(n, dt), np.full(n, mx, dt)]


@pytest.mark.parametrize("depth", [8, 10])
@pytest.mark.parametrize("kind,fn", [(0, "ola_sad8x8"), (1, "ola_satd8x8"), (2, "ola_sa8d8x8"), (3, "ola_sa8d16x16")])
def test_pixelcmp_batch(mods, depth, kind, fn):
    """check_pixelcmp recipe (pixelharness.cpp:80-99): stride 64 vs FENC_STRIDE-5-like odd stride,
    offsets stepping through the buffers, all combinations of random/min/max buffers"""
    replay, po, abi = mods
    lib = po.oracle(depth)
    bufs = _pixel_buffers(po, depth, 1234 + kind)
    la = abi.Lookahead(64, 64, depth, 1, 4, 0, 0, 0, 0, 0.0)
    try:
        for ia, a in enumerate(bufs):
            for ib, b in enumerate(bufs):
                sa, sb = 64, 59
                n = 100
                offA = (np.arange(n, dtype=np.int64) * 37) % 1500
                offB = (np.arange(n, dtype=np.int64) * 53 + 3) % 1500
                out = np.zeros(n, np.int32)
                r = abi.lib_cu().x265cu_pixelcmp_batch(la.ctx, kind, a.ctypes.data, a.size, sa, b.ctypes.data, b.size, sb, n,
                                                       offA.ctypes.data, offB.ctypes.data, out.ctypes.data)
                assert r == 0, abi.lib_cu().x265cu_last_error(la.ctx)
                isz = a.itemsize
                want = np.array([getattr(lib, fn)(a.ctypes.data + int(oa) * isz, sa, b.ctypes.data + int(ob) * isz, sb)
                                 for oa, ob in zip(offA, offB)], np.int32)
                assert np.array_equal(out, want), (ia, ib, out[:8], want[:8])
    finally:
        la.close()


PU_SHAPES = [(4, 4), (8, 8), (16, 16), (32, 32), (64, 64), (8, 4), (4, 8), (16, 8), (8, 16), (32, 16), (16, 32), (64, 32), (32, 64), (16, 12),
             (12, 16), (16, 4), (4, 16), (32, 24), (24, 32), (32, 8), (8, 32), (64, 48), (48, 64), (64, 16), (16, 64)]


@pytest.mark.parametrize("depth", [8, 10])
@pytest.mark.parametrize("kind,fn", [(0, "ola_pu_sad"), (1, "ola_pu_satd")])
def test_pixelcmp_pu_shapes(mods, depth, kind, fn):
    """pu[LUMA_WxH].sad / .satd of the 25 luma PU shapes (common/pixel.cpp:954-1004) as a batch operation ==
    the oracle (pinned against the reference's own primitives table per shape in tests/test_oracle_vs_ref.py); TestBench
    buffers (random / min / max), FENC_STRIDE-like and odd strides, unaligned offsets"""
    replay, po, abi = mods
    lib = po.oracle(depth)
    f = getattr(lib, fn)
    f.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_ssize_t, C.c_void_p, C.c_ssize_t]
    L = abi.lib_cu()
    L.x265cu_pixelcmp_pu.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_ssize_t, C.c_void_p, C.c_size_t, C.c_ssize_t,
                                     C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    bufs = _pixel_buffers(po, depth, 4321 + kind)
    la = abi.Lookahead(64, 64, depth, 1, 4, 0, 0, 0, 0, 0.0)
    try:
        for w, h in PU_SHAPES:
            for a, b in ((bufs[0], bufs[0][::-1].copy()), (bufs[0], bufs[1]), (bufs[2], bufs[0]), (bufs[1], bufs[2])):
                sa, sb = 64, 131
                n = 40
                lim_a, lim_b = a.size - (h - 1) * sa - w, b.size - (h - 1) * sb - w
                offA = (np.arange(n, dtype=np.int64) * 37) % lim_a
                offB = (np.arange(n, dtype=np.int64) * 53 + 3) % lim_b
                out = np.zeros(n, np.int32)
                r = L.x265cu_pixelcmp_pu(la.ctx, kind, w, h, a.ctypes.data, a.size, sa, b.ctypes.data, b.size, sb, n, offA.ctypes.data, offB.ctypes.data, out.ctypes.data)
                assert r == 0, L.x265cu_last_error(la.ctx)
                isz = a.itemsize
                want = np.array([f(w, h, a.ctypes.data + int(oa) * isz, sa, b.ctypes.data + int(ob) * isz, sb) for oa, ob in zip(offA, offB)], np.int32)
                assert np.array_equal(out, want), (w, h, out[:6], want[:6])
        # not a PU shape / a block outside its buffer: refused, not computed
        out = np.zeros(1, np.int32)
        off = np.zeros(1, np.int64)
        a = bufs[0]
        assert L.x265cu_pixelcmp_pu(la.ctx, kind, 24, 24, a.ctypes.data, a.size, 64, a.ctypes.data, a.size, 64, 1, off.ctypes.data, off.ctypes.data, out.ctypes.data) == -1
        far = np.array([a.size - 10], np.int64)
        assert L.x265cu_pixelcmp_pu(la.ctx, kind, 8, 8, a.ctypes.data, a.size, 64, a.ctypes.data, a.size, 64, 1, far.ctypes.data, off.ctypes.data, out.ctypes.data) == -1
        assert L.x265cu_pixelcmp_batch(la.ctx, 0, a.ctypes.data, a.size, 64, a.ctypes.data, a.size, 64, 1, far.ctypes.data, off.ctypes.data, out.ctypes.data) == -1
    finally:
        la.close()


def test_bad_arguments_fail_loudly(mods):
    replay, po, abi = mods
    L = abi.lib_cu()
    la = abi.Lookahead(320, 192, 8, 3, 10, 0, 0, 1, 1, 1.0)
    try:
        assert L.x265cu_frame_init(la.ctx, 9999, None, 0, 0, None) == -1
        res = (abi.JobResult * 1)()
        job = (abi.Job * 1)()
        job[0].fenc = 0; job[0].ref0 = 0; job[0].d0 = 0
        assert L.x265cu_estimate_batch(la.ctx, 1, job, res) == -1
        assert b"bad job" in L.x265cu_last_error(la.ctx)
    finally:
        la.close()


@pytest.mark.parametrize("wide", ["0", "1", "2", "1 looping"])
@pytest.mark.parametrize("name", ["odd8", "tiny10"])
def test_pixelcmp_frames(mods, monkeypatch, name, wide):
    """the device-resident whole-frame form of the primitives (the "SATD Gpix/s" kernel of the bench): SAD / SATD / SA8D
    of every aligned 8x8 block of frame pairs, all three kernel forms (quad, wide x1, wide x2), odd CU counts (odd8:
    23 x 13 CUs), against the oracle on the planes the GPU itself produced (those are CRC-checked by the replay)"""
    replay, po, abi = mods
    monkeypatch.setenv("X265CU_PIXELCMP_WIDE", wide.split()[0])
    if "looping" in wide:
        monkeypatch.setenv("X265CU_PIXELCMP_BX", "1")      # one CTA per pair: its warps loop over the whole plane
    t = po.Trace(replay.trace_path(name))
    r = replay.CuReplay(t, check=False)
    try:
        r.run(max_events=6)
        la, L = r.la, abi.lib_cu()
        la.sync()
        pocs = sorted(r.frames)[:4]
        lib = po.oracle(t.cfg["depth"])
        dt = po.pixel_dtype(t.cfg["depth"])
        planes = {p: la.array(r.frames[p], 0, dtype=dt).copy() for p in pocs}
        slots = {p: abi.lib_host().x265cuh_frame_slot(r.frames[p]) for p in pocs}
        pairs = [(pocs[0], pocs[1]), (pocs[2], pocs[1]), (pocs[3], pocs[3]), (pocs[0], pocs[3])]
        a = np.array([slots[x] for x, _ in pairs], np.int32)
        b = np.array([slots[y] for _, y in pairs], np.int32)
        pad = la.padOffset
        for kind, fn in ((0, "ola_sad8x8"), (1, "ola_satd8x8"), (2, "ola_sa8d8x8")):
            out = np.zeros(len(pairs) * la.nCU, np.int32)
            assert L.x265cu_pixelcmp_frames(la.ctx, kind, len(pairs), a.ctypes.data, b.ctypes.data, out.ctypes.data, None) == 0
            for k, (x, y) in enumerate(pairs):
                pa, pb = planes[x], planes[y]
                want = np.array([getattr(lib, fn)(pa.ctypes.data + (pad + 8 * (cu // la.wCU) * la.stride + 8 * (cu % la.wCU)) * pa.itemsize, la.stride,
                                                  pb.ctypes.data + (pad + 8 * (cu // la.wCU) * la.stride + 8 * (cu % la.wCU)) * pb.itemsize, la.stride)
                                 for cu in range(la.nCU)], np.int32)
                got = out[k * la.nCU:(k + 1) * la.nCU]
                assert np.array_equal(got, want), (name, wide, kind, k, np.flatnonzero(got != want)[:6])
        # plane selection: half-pel planes H (1) against C (3) of the first pair
        pA, pB = np.array([1], np.int32), np.array([3], np.int32)
        out = np.zeros(la.nCU, np.int32)
        assert L.x265cu_pixelcmp_planes(la.ctx, 1, 1, a.ctypes.data, pA.ctypes.data, b.ctypes.data, pB.ctypes.data, out.ctypes.data, None) == 0
        x, y = pairs[0]
        offA, offB = la.planeSize * 1, la.planeSize * 3
        want = np.array([lib.ola_satd8x8(planes[x].ctypes.data + (offA + pad + 8 * (cu // la.wCU) * la.stride + 8 * (cu % la.wCU)) * planes[x].itemsize, la.stride,
                                         planes[y].ctypes.data + (offB + pad + 8 * (cu // la.wCU) * la.stride + 8 * (cu % la.wCU)) * planes[y].itemsize, la.stride)
                         for cu in range(la.nCU)], np.int32)
        assert np.array_equal(out, want), (name, wide, "planes")
        assert L.x265cu_pixelcmp_planes(la.ctx, 1, 1, a.ctypes.data, np.array([4], np.int32).ctypes.data, b.ctypes.data, None, None, None) == -1
    finally:
        r.close()
