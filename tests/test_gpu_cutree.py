"""cuTree propagation on the GPU against the oracle, through the C ABI (x265cu_cutree_run).

The golden traces (tests/test_gpu_parity.py) cover every propagate step the reference ran on the clips; here the
same entry point is driven with adversarial arrays the clips never produce: saturating propagateCost, vectors that
leave the picture on every side, intra CUs, bipred weights, inverse qscales 16x the neutral one, lists of more ops
than one launch holds.  (Inputs stay inside the domain where the reference's own int32 arithmetic does not overflow,
listamount * 1024 < 2^31; beyond it the reference is undefined behaviour.  The wrapping intraCost * invQscale
product of the primitive alone is pinned in tests/test_oracle_vs_ref.py.)  The oracle's ola_estimate_cu_propagate (pinned against the reference on all traces and,
for its primitive, against the reference's own compiled function) is the checker."""
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


class Rig:
    """nslots frames of a small picture on the device (real lowres planes + intra costs) and their oracle twins"""

    def __init__(self, po, abi, depth, w, h, bframes, nslots, seed):
        self.po, self.abi, self.depth, self.bf = po, abi, depth, bframes
        self.L = abi.lib_cu()
        self.la = abi.Lookahead(w, h, depth, bframes, 10, 0, 0, 0, 1, 1.0, slots=nslots)
        self.ctx = self.la.ctx
        self.lib = po.oracle(depth)
        self.n = self.la.nCU
        self.rng = np.random.default_rng(seed)
        self.of = []
        for s in range(nslots):
            f = po.OFrame(self.lib, depth, w, h, 96, 80, bframes, True)
            y, u, v = po.synth_padded(self.lib, depth, w, h, s, nslots, seed)
            assert self.L.x265cu_frame_init(self.ctx, s, y.ctypes.data, y.strides[0] // y.itemsize, 0, None) == 0
            out = abi.IntraOut()
            ic = f.intra_cost()
            out.intraCost = ic.ctypes.data
            assert self.L.x265cu_intra(self.ctx, s, C.byref(out)) == 0      # intraCost now on the device and in the twin
            self.of.append(f)

    def close(self):
        for f in self.of:
            f.close()
        self.la.close()

    def set_invq(self, s, a):
        a = np.ascontiguousarray(a, np.int32)
        self.of[s].inv_qscale()[:] = a
        assert self.L.x265cu_frame_set_invqscale(self.ctx, s, a.ctypes.data) == 0
        assert self.L.x265cu_sync(self.ctx) == 0

    def set_costs(self, s, d0, d1, a):
        a = np.ascontiguousarray(a, np.uint16)
        self.of[s].lowres_costs(d0, d1)[:] = a
        assert self.L.x265cu_frame_set_array(self.ctx, s, 4, d0, d1, a.ctypes.data) == 0

    def set_mvs(self, s, lst, d, a):
        a = np.ascontiguousarray(a, np.int16)
        self.of[s].mvs(lst, d)[:] = a
        assert self.L.x265cu_frame_set_array(self.ctx, s, 6, lst, d, a.ctypes.data) == 0

    def set_prop(self, s, a):
        a = np.ascontiguousarray(a, np.uint16)
        self.of[s].propagate_cost()[:] = a
        assert self.L.x265cu_frame_set_propagate(self.ctx, s, a.ctypes.data) == 0

    def run(self, ops, fetch):
        """ops: ("Z", slot) | ("X", b, p0, p1, referenced, bipredWeight(None = no weighted bipred), fpsFactor)"""
        abi = self.abi
        arr = (abi.CutreeOp * max(len(ops), 1))()
        for i, o in enumerate(ops):
            if o[0] == "Z":
                arr[i].kind = 0
                arr[i].fenc = o[1]
                self.lib.ola_cutree_zero(self.of[o[1]].p)
            else:
                _, b, p0, p1, referenced, weighted, fps = o
                d0, d1 = b - p0, p1 - b
                dsf = ((d0 << 8) + ((p1 - p0) >> 1)) // (p1 - p0)
                arr[i].kind = 1
                arr[i].fenc, arr[i].ref0, arr[i].ref1, arr[i].d0, arr[i].d1 = b, p0, p1, d0, d1
                arr[i].referenced = referenced
                arr[i].bipredWeight = 64 - (dsf >> 2) if weighted else 32
                arr[i].fpsFactor = fps
                # oracle: fpsFactor = clip(fpsDenom / fpsNum) / clip(avgDuration); choose avgDuration = 1 / 30 / fps * ... simpler:
                # drive it with fpsNum/fpsDenom = 30/1 and averageDuration = (1/30) / fps, both inside the clip range
                self.lib.ola_estimate_cu_propagate(self.of[b].p, self.of[p0].p, self.of[p1].p, d0, d1, referenced,
                                                   (1.0 / 30) / fps, 30, 1, 1 if weighted else 0)
        outs = [np.zeros(self.n, np.uint16) for _ in fetch]
        slots = (C.c_int * max(len(fetch), 1))(*fetch)
        ptrs = (C.c_void_p * max(len(fetch), 1))(*[o.ctypes.data for o in outs])
        r = self.L.x265cu_cutree_run(self.ctx, len(ops), arr, len(fetch), slots, ptrs)
        assert r == 0, self.L.x265cu_last_error(self.ctx)
        return outs


def _fps_exact(fps):
    """the oracle recomputes fpsFactor from durations; only use factors that survive that round trip exactly"""
    return (1.0 / 30) / ((1.0 / 30) / fps) == fps


@pytest.mark.parametrize("depth", [8, 10])
def test_propagate_adversarial(mods, depth):
    po, abi = mods
    w, h, bf = 352, 208, 3           # lowres 176x104 -> 22x13 CUs
    rig = Rig(po, abi, depth, w, h, bf, 5, 77 + depth)
    try:
        n, rng = rig.n, rig.rng
        wCU = rig.la.wCU
        assert _fps_exact(1.0)
        for rnd in range(6):
            for s in range(5):
                invq = rng.integers(16, 2048, n)
                if rnd == 3:
                    invq[: n // 3] = 4096              # 16x the neutral factor: amounts around 2^20
                rig.set_invq(s, invq)
                # propagateCost near and at the saturation point
                rig.set_prop(s, rng.choice([0, 1, 7, 40000, 65000, 65534, 65535], n) if rnd % 2 else rng.integers(0, 65536, n))
            b, p0, p1 = 2, 0, 4
            # lowresCosts: cost in the low 14 bits (some above the intra cost), lists in the top 2
            ic = rig.of[b].intra_cost()
            cost = np.minimum(rng.integers(0, 1 << 14, n), np.where(rng.random(n) < 0.7, ic // 2, 1 << 14)).astype(np.int64) & 0x3FFF
            lists = rng.integers(0, 4, n)
            rig.set_costs(b, 2, 2, cost | (lists << 14))
            rig.set_costs(b, 2, 0, cost | (rng.integers(0, 2, n) << 14))
            for lst, d in ((0, 2), (1, 2)):
                mv = rng.integers(-64, 65, (n, 2))
                big = rng.random(n) < 0.3
                mv[big] = rng.integers(-4000, 4001, (int(big.sum()), 2))     # far outside the picture
                mv[rng.random(n) < 0.2] = 0                                   # the mv0 shortcut
                mv[:wCU, 1] = rng.integers(-40, 1, wCU)                       # top row pointing up / partly outside
                rig.set_mvs(b, lst, d, mv.reshape(-1))
            ops = [("X", b, p0, p1, 1, rnd & 1, 1.0),             # bidir, referenced
                   ("X", b, p0, p1, 0, 0, 1.0),                   # same, non-referenced: zero row trick
                   ("Z", 1),
                   ("X", 2, 0, 2, 1, 0, 1.0),                     # P-type (d1 = 0)
                   ("X", b, p0, p1, 1, 1, 1.0)]
            got = rig.run(ops, [0, 1, 2, 4])
            for s, g in zip([0, 1, 2, 4], got):
                want = rig.of[s].propagate_cost()
                assert np.array_equal(g, want), (depth, rnd, s, np.flatnonzero(g != want)[:8], g[g != want][:8], want[g != want][:8])
    finally:
        rig.close()


def test_long_op_list_spans_launches(mods):
    """more ops than one launch holds (CUTREE_MAX_OPS = 64): the list continues in the next launch, in order"""
    po, abi = mods
    rig = Rig(po, abi, 8, 352, 208, 3, 5, 5)
    try:
        n, rng = rig.n, rig.rng
        for s in range(5):
            rig.set_invq(s, rng.integers(200, 400, n))
            rig.set_prop(s, np.zeros(n))
        ic = rig.of[2].intra_cost()
        for (d0, d1) in ((2, 2), (1, 1), (2, 0), (1, 0)):
            for s in range(5):
                rig.set_costs(s, d0, d1, (rig.of[s].intra_cost() // 3 & 0x3FFF) | ((3 if d1 else 1) << 14))
        for s in range(5):
            for lst in (0, 1):
                for d in (1, 2):
                    rig.set_mvs(s, lst, d, rng.integers(-48, 49, 2 * n))
        ops = []
        for k in range(60):
            ops += [("X", 2, 0, 4, 1, 0, 1.0), ("X", 1, 0, 2, 0, 0, 1.0), ("X", 3, 2, 4, 0, 0, 1.0), ("X", 4, 2, 4, 1, 0, 1.0)]
            if k % 7 == 6:
                ops.append(("Z", 4))
        assert len(ops) > 160
        got = rig.run(ops, [0, 2, 4, 1, 3])
        for s, g in zip([0, 2, 4, 1, 3], got):
            assert np.array_equal(g, rig.of[s].propagate_cost()), s
        assert got[0].max() == 65535     # the chain saturated somewhere
    finally:
        rig.close()


def test_cutree_bad_arguments(mods):
    po, abi = mods
    L = abi.lib_cu()
    la = abi.Lookahead(320, 192, 8, 3, 10, 0, 0, 1, 1, 1.0)
    try:
        op = (abi.CutreeOp * 1)()
        op[0].kind = 1; op[0].fenc = 0; op[0].ref0 = 0; op[0].ref1 = 9999; op[0].d0 = 1
        assert L.x265cu_cutree_run(la.ctx, 1, op, 0, None, None) == -1
        op[0].ref1 = 0; op[0].d0 = 0
        assert L.x265cu_cutree_run(la.ctx, 1, op, 0, None, None) == -1
        op[0].d0 = 1           # frame 0 never got an invQscaleFactor
        assert L.x265cu_cutree_run(la.ctx, 1, op, 0, None, None) == -1
        assert b"invQscaleFactor" in L.x265cu_last_error(la.ctx)
        assert L.x265cu_frame_set_array(la.ctx, 0, 5, 0, 0, op) == -1
        assert L.x265cu_frame_set_propagate(la.ctx, 0, None) == -1
    finally:
        la.close()
