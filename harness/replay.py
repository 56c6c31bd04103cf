"""Replay of a reference lookahead TRACE through the product (host layer -> C ABI -> CUDA kernels).

A trace (tests/golden/*.trace, produced by the UNMODIFIED reference through oracle/ref_shim.cpp)
lists every pre-lookahead frame and every frame-cost estimate the reference's lookahead ran on a
deterministic synthetic clip, in order, with CRC32s of all output arrays.  Replaying it through
x265cu::Lookahead / CostEstimateGroup reproduces the exact call pattern of the real host
(batches via add()/finishBatch(), the rest via singleCost()) and lets every output be compared
bit-for-bit at full size without the reference being present.

This module is test/bench harness code: it may use the oracle (synthetic clip generator, trace
parser); the product package src/x265_b200 never imports it.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import pyoracle as po  # noqa: E402
from src.x265_b200 import abi  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def trace_path(name):
    return os.path.join(GOLD, name + ".trace")


class Clip:
    """Synthetic clip of a trace, generated once into (pinned-size) host arrays padded like PicYuv."""

    def __init__(self, cfg, frames=None):
        self.cfg = cfg
        lib = po.oracle(cfg["depth"])
        self.frames = {}
        want = range(cfg["nframes"]) if frames is None else frames
        for t in want:
            self.frames[t] = po.synth_padded(lib, cfg["depth"], cfg["width"], cfg["height"], t, cfg["nframes"], cfg["seed"])


class CuReplay:
    def __init__(self, trace, device=0, check=True, slots=0, clip=None, planes_back=True):
        self.t = trace
        cfg = trace.cfg
        self.cfg = cfg
        # lookaheadSlices as the user gave it is not in the trace; numCoopSlices is: reproduce it
        slices = cfg["numCoopSlices"] if cfg["numCoopSlices"] > 1 else 0
        ct = dict(fps_num=cfg.get("fpsNum", 30), fps_denom=cfg.get("fpsDenom", 1), qcompress=cfg.get("qCompress", 0.6),
                  weighted_bipred=cfg.get("weightedBiPred", 0))
        self.la = abi.Lookahead(cfg["width"], cfg["height"], cfg["depth"], cfg["bframes"], cfg["lookahead"], slices,
                                cfg["pool"], cfg["weightp"], cfg["aqmode"], cfg["aqStrength"], cfg["bFrameBias"], device, slots, **ct)
        if slices and (self.la.numCoopSlices, self.la.numRowsPerSlice) != (cfg["numCoopSlices"], cfg["numRowsPerSlice"]):
            # numCoopSlices = H8 / rowsPerSlice is not invertible in general; fall back to a search
            self.la.close()
            for s in range(2, 17):
                self.la = abi.Lookahead(cfg["width"], cfg["height"], cfg["depth"], cfg["bframes"], cfg["lookahead"], s,
                                        cfg["pool"], cfg["weightp"], cfg["aqmode"], cfg["aqStrength"], cfg["bFrameBias"], device, slots, **ct)
                if (self.la.numCoopSlices, self.la.numRowsPerSlice) == (cfg["numCoopSlices"], cfg["numRowsPerSlice"]):
                    break
                self.la.close()
            else:
                raise RuntimeError("cannot reproduce the trace's slice geometry")
        self.check = check
        self.clip = clip
        self.planes_back = planes_back
        self.frames = {}        # poc -> handle
        self.free = []
        self.keep = cfg["lookahead"] + cfg["bframes"] + 4
        self.mismatches = []
        self.njobs = 0
        self.nframes = 0
        self.npropagate = 0
        self.cutree = True      # replay the cuTree events (M/X/F) of the trace
        self.lib = po.oracle(cfg["depth"])

    def close(self):
        for f in list(self.frames.values()) + self.free:
            self.la.frame_free(f)
        self.frames = {}
        self.free = []
        self.la.close()

    def _chk(self, what, got, want):
        if got != want:
            self.mismatches.append((what, got, want))

    def pre(self, e):
        self.pre_list([e])

    def pre_list(self, es):
        """one PreLookaheadGroup list: a single frame goes through preLookahead(), several through preLookaheadBatch()"""
        cfg = self.cfg
        items, keep = [], []
        for e in es:
            poc = e["poc"]
            for old in [p for p in self.frames if p < min(x["poc"] for x in es) - self.keep]:
                self.free.append(self.frames.pop(old))
            f = self.free.pop() if self.free else self.la.frame_alloc()
            if self.clip is not None:
                y, u, v = self.clip.frames[poc]
            else:
                y, u, v = po.synth_padded(self.lib, cfg["depth"], cfg["width"], cfg["height"], poc, cfg["nframes"], cfg["seed"])
            keep.append((y, u, v))
            items.append((f, y.ctypes.data, y.strides[0] // y.itemsize, u.ctypes.data, v.ctypes.data, u.strides[0] // u.itemsize, poc))
            self.frames[poc] = f
            self.nframes += 1
        if len(es) == 1:
            self.la.pre_lookahead(items[0][0], keep[0][0], keep[0][1], keep[0][2], es[0]["poc"], self.planes_back)
        else:
            prep = self.la.prepare_pre_lookahead_batch(items)
            # Lookahead::addPicture for every other frame of the list: the pre-lookahead must take pictures uploaded ahead
            # and pictures it uploads itself side by side
            ahead = self.la.prepare_pre_lookahead_batch(items[::2])
            self.la.add_pictures_prepared(ahead)
            self.la.pre_lookahead_batch_prepared(prep, self.planes_back)
        if not self.check:
            return
        for e, it in zip(es, items):
            self._check_pre(e, it[0])

    def _check_pre(self, e, f):
        cfg = self.cfg
        poc = e["poc"]
        la, tag = self.la, "P%d." % poc
        if self.planes_back:
            la.sync()      # the padded planes come back behind the compute stream
            self._chk(tag + "planes", la.crc(f, 0), e["planes"])
        if cfg["aqmode"]:
            self._chk(tag + "invQ", la.crc(f, 3), e["invQ"])
        s = la.scalars(f, 0, 0)
        self._chk(tag + "wp_ssd", s["wp_ssd"], e["wp_ssd"])
        self._chk(tag + "wp_sum", s["wp_sum"], e["wp_sum"])
        self._chk(tag + "intraCost", la.crc(f, 1), e["intraCost"])
        self._chk(tag + "intraMode", la.crc(f, 2), e["intraMode"])
        self._chk(tag + "lowresCosts", la.crc(f, 4, 0, 0), e["lowresCosts"])
        self._chk(tag + "rowSatds", la.crc(f, 5, 0, 0), e["rowSatds"])
        self._chk(tag + "costEst", s["costEst"], e["costEst"])
        self._chk(tag + "costEstAq", s["costEstAq"], e["costEstAq"])

    def _run_jobs(self, jobs, batch):
        # like the reference's frames[] array, index distance must equal POC distance
        lo = min(j["p0"] for j in jobs)
        hi = max(j["p1"] for j in jobs)
        frames = [self.frames.get(p) for p in range(lo, hi + 1)]
        triples = [(j["p0"] - lo, j["p1"] - lo, j["b"] - lo) for j in jobs]
        scores = self.la.estimate(frames, triples, batch)
        self.njobs += len(jobs)
        if not self.check:
            return
        la = self.la
        for j, score in zip(jobs, scores):
            f = self.frames[j["b"]]
            d0, d1 = j["b"] - j["p0"], j["p1"] - j["b"]
            tag = "J%d/%d/%d." % (j["p0"], j["b"], j["p1"])
            want = j["costEst"] * 100 // (130 + self.cfg["bFrameBias"]) if d1 > 0 else j["costEst"]
            s = la.scalars(f, d0, d1)
            if j["s0"]:
                self._chk(tag + "weight", (s["weighted"], s["wscale"], s["wdenom"], s["woffset"]) if s["weighted"] else (0, 0, 0, 0),
                          (1, j["wscale"], j["wdenom"], j["woffset"]) if j["wflag"] == 1 else (0, 0, 0, 0))
            self._chk(tag + "costEst", score, want)
            self._chk(tag + "costEstAq", s["costEstAq"], j["costEstAq"])
            if d1 == 0:
                self._chk(tag + "intraMbs", s["intraMbs"], j["intraMbs"])
            if d0 > 0:
                self._chk(tag + "mv0", la.crc(f, 6, 0, d0), j["mv0"])
                self._chk(tag + "mc0", la.crc(f, 7, 0, d0), j["mc0"])
            if d1 > 0:
                self._chk(tag + "mv1", la.crc(f, 6, 1, d1), j["mv1"])
                self._chk(tag + "mc1", la.crc(f, 7, 1, d1), j["mc1"])
            self._chk(tag + "lowresCosts", la.crc(f, 4, d0, d1), j["lowresCosts"])
            self._chk(tag + "rowSatds", la.crc(f, 5, d0, d1), j["rowSatds"])

    # ---- cuTree: the memsets, propagate steps and cuTreeFinish calls of Lookahead::cuTree, in the
    # reference's order, interleaved with its estimates exactly as the trace recorded them
    def ct_propagate(self, x):
        lo, hi = x["p0"], x["p1"]
        frames = [self.frames.get(p) for p in range(lo, hi + 1)]
        self.la.cutree_propagate(frames, 0, hi - lo, x["b"] - lo, x["referenced"], x["avgDuration"])
        self.npropagate += 1
        if not self.check:
            return
        la, tag = self.la, "X%d/%d/%d." % (x["p0"], x["b"], x["p1"])
        self._chk(tag + "ref0", la.crc(self.frames[x["p0"]], 8), x["ref0"])
        if x["p1"] != x["b"]:
            self._chk(tag + "ref1", la.crc(self.frames[x["p1"]], 8), x["ref1"])
        self._chk(tag + "own", la.crc(self.frames[x["b"]], 8, 0 if x["referenced"] else 1), x["own"])

    def ct_finish(self, e):
        f = self.frames[e["poc"]]
        self.la.cutree_finish(f, e["avgDuration"], e["ref0Distance"])
        if not self.check:
            return
        self._chk("F%d.propagateCost" % e["poc"], self.la.crc(f, 8), e["propagateCost"])
        self._chk("F%d.qpCuTreeOffset" % e["poc"], self.la.crc(f, 9), e["qpCuTreeOffset"])

    def run(self, max_events=None, stop_on_mismatch=False):
        n = 0
        pend = []
        for e in self.t.events:
            if e[0] == "P":
                pend.append(e[1])     # consecutive frames form one PreLookaheadGroup list
                n += 1
                if not (max_events and n >= max_events):
                    continue
            if pend:
                self.pre_list(pend)
                pend = []
            if e[0] == "P":
                break
            if e[0] == "J":
                self._run_jobs([e[1]], False)
            elif e[0] == "B" and e[1]:
                self._run_jobs(e[1], True)
            elif self.cutree and e[0] == "M":
                self.la.cutree_zero(self.frames[e[1]])
            elif self.cutree and e[0] == "X":
                self.ct_propagate(e[1])
            elif self.cutree and e[0] == "F":
                self.ct_finish(e[1])
            n += 1
            if max_events and n >= max_events:
                break
            if stop_on_mismatch and self.mismatches:
                break
        if pend:
            self.pre_list(pend)
        return self.mismatches


def replay_trace(name, device=0, max_events=None, check=True):
    t = po.Trace(trace_path(name))
    r = CuReplay(t, device=device, check=check)
    try:
        t0 = time.time()
        mm = r.run(max_events=max_events)
        return dict(name=name, jobs=r.njobs, frames=r.nframes, propagates=r.npropagate, cutree=r.la.cutree_stats(),
                    seconds=time.time() - t0, mismatches=mm)
    finally:
        r.close()


def smoke():
    """tiny8 trace on cuda:0, every CRC compared with the reference's; plus one direct oracle check
    of the SATD primitive through the C ABI."""
    res = replay_trace("tiny8")
    if res["mismatches"]:
        return False, "tiny8 replay: %d mismatches, first %r" % (len(res["mismatches"]), res["mismatches"][0])
    # primitive check against the oracle
    import ctypes as C
    lib = po.oracle(8)
    rng = np.random.default_rng(1)
    a = rng.integers(0, 256, 64 * 64, dtype=np.uint8)
    b = rng.integers(0, 256, 64 * 64, dtype=np.uint8)
    offs = np.arange(0, 40, dtype=np.int64) * 65 % (64 * 56 - 8)
    out = np.zeros(40, np.int32)
    la = abi.Lookahead(64, 64, 8, 1, 4, 0, 0, 0, 0, 0.0)
    try:
        r = abi.lib_cu().x265cu_pixelcmp_batch(la.ctx, 1, a.ctypes.data, a.size, 64, b.ctypes.data, b.size, 64, 40,
                                               offs.ctypes.data, offs.ctypes.data, out.ctypes.data)
        if r:
            return False, "pixelcmp_batch failed %d" % r
    finally:
        la.close()
    for i, o in enumerate(offs):
        want = lib.ola_satd8x8(a.ctypes.data + int(o), 64, b.ctypes.data + int(o), 64)
        if want != out[i]:
            return False, "satd mismatch at %d: %d vs %d" % (i, out[i], want)
    return True, "tiny8 replay (%d frames, %d estimates, %d cuTree propagate steps) bit-exact vs reference trace; SATD batch == oracle" % (
        res["frames"], res["jobs"], res["propagates"])


if __name__ == "__main__":
    names = sys.argv[1:] or ["tiny8"]
    for nm in names:
        r = replay_trace(nm)
        print("%-12s frames %d jobs %d propagates %d %r  %.2fs  mismatches %d" % (nm, r["frames"], r["jobs"], r["propagates"], r["cutree"],
                                                                                r["seconds"], len(r["mismatches"])))
        for m in r["mismatches"][:12]:
            print("    ", m)
