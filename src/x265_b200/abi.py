"""ctypes bindings of the product libraries: libx265cu.so (C ABI, include/x265cu.h) and
libx265cu_host.so (C++ host layer, host/lookahead_cu.h, through its flat x265cuh_* view).

There is no fallback of any kind: if the libraries are missing, loading raises; if there is no
CUDA device, x265cu_open fails with X265CU_ENODEV and `Lookahead` raises.
"""
import ctypes as C
import os

import numpy as np

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_CU = os.path.join(PKG, "libx265cu.so")
LIB_HOST = os.path.join(PKG, "libx265cu_host.so")

BFMAX = 16
K_NAMES = ("lowres", "intra", "search", "cost", "weight", "pixel", "var", "cutree", "results")


class Config(C.Structure):
    _fields_ = [("srcWidth", C.c_int), ("srcHeight", C.c_int), ("bitDepth", C.c_int), ("marginX", C.c_int), ("marginY", C.c_int),
                ("bframes", C.c_int), ("numFrameSlots", C.c_int), ("numCoopSlices", C.c_int), ("numRowsPerSlice", C.c_int),
                ("bFrameBias", C.c_int), ("lookaheadLambda", C.c_int), ("mvcost", C.c_void_p), ("device", C.c_int),
                ("stream", C.c_void_p), ("searchWarps", C.c_int)]


class Geometry(C.Structure):
    _fields_ = [("width", C.c_int), ("lines", C.c_int), ("stride", C.c_int), ("paddedLines", C.c_int),
                ("widthInCU", C.c_int), ("heightInCU", C.c_int), ("cuCount", C.c_int),
                ("planeSize", C.c_int64), ("padOffset", C.c_int64), ("pixelBytes", C.c_int)]


class IntraOut(C.Structure):
    _fields_ = [("intraCost", C.c_void_p), ("intraMode", C.c_void_p), ("lowresCosts", C.c_void_p), ("rowSatds", C.c_void_p),
                ("sums", C.c_int64 * 2)]


class WeightItem(C.Structure):
    _fields_ = [("fenc", C.c_int), ("ref", C.c_int), ("weighted", C.c_int), ("scale", C.c_int), ("denom", C.c_int), ("offset", C.c_int)]


class Job(C.Structure):
    _fields_ = [("fenc", C.c_int), ("ref0", C.c_int), ("ref1", C.c_int), ("d0", C.c_int), ("d1", C.c_int),
                ("doSearch", C.c_int * 2), ("sliced", C.c_int), ("weighted", C.c_int),
                ("wScale", C.c_int), ("wDenom", C.c_int), ("wOffset", C.c_int),
                ("mvs", C.c_void_p * 2), ("mvCosts", C.c_void_p * 2), ("lowresCosts", C.c_void_p), ("rowSatds", C.c_void_p)]


class JobResult(C.Structure):
    _fields_ = [("costEstRaw", C.c_int64), ("costEst", C.c_int64), ("costEstAq", C.c_int64), ("intraMbs", C.c_int32), ("reserved", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [("ms", C.c_double * 9), ("launches", C.c_int64 * 9), ("h2dBytes", C.c_int64), ("d2hBytes", C.c_int64)]


class HostParams(C.Structure):
    _fields_ = [("sourceWidth", C.c_int), ("sourceHeight", C.c_int), ("bitDepth", C.c_int), ("maxCUSize", C.c_int),
                ("bframes", C.c_int), ("lookaheadDepth", C.c_int), ("lookaheadSlices", C.c_int), ("poolWorkers", C.c_int),
                ("bEnableWeightedPred", C.c_int), ("aqMode", C.c_int), ("aqStrength", C.c_double),
                ("bFrameBias", C.c_int), ("device", C.c_int), ("frameSlots", C.c_int), ("stream", C.c_void_p), ("searchWarps", C.c_int),
                ("fpsNum", C.c_int), ("fpsDenom", C.c_int), ("qCompress", C.c_double), ("bEnableWeightedBiPred", C.c_int)]


class CutreeOp(C.Structure):
    _fields_ = [("kind", C.c_int), ("fenc", C.c_int), ("ref0", C.c_int), ("ref1", C.c_int), ("d0", C.c_int), ("d1", C.c_int),
                ("referenced", C.c_int), ("bipredWeight", C.c_int), ("fpsFactor", C.c_double)]


# every symbol include/x265cu.h declares (tests check that the library exports all of them)
ABI_SYMBOLS = (
    "x265cu_abi_version", "x265cu_device_count", "x265cu_open", "x265cu_close", "x265cu_last_error",
    "x265cu_get_geometry", "x265cu_sync", "x265cu_host_register", "x265cu_host_unregister", "x265cu_trim", "x265cu_wp_prepare", "x265cu_wp_cost", "x265cu_pixelcmp_pu", "x265cu_motion_estimate",
    "x265cu_frame_init", "x265cu_frame_set_invqscale", "x265cu_frame_var", "x265cu_frame_init_var", "x265cu_frame_init_var_batch",
    "x265cu_intra", "x265cu_intra_batch", "x265cu_pre_lookahead_batch", "x265cu_frame_upload",
    "x265cu_weight_cost_batch", "x265cu_estimate_batch", "x265cu_pixelcmp_batch", "x265cu_pixelcmp_frames", "x265cu_pixelcmp_planes", "x265cu_int_peak",
    "x265cu_stats_enable", "x265cu_stats_get",
    "x265cu_cutree_run", "x265cu_frame_set_propagate", "x265cu_frame_set_array",
)

_cu = None
_host = None


def lib_cu():
    """libx265cu.so; raises OSError when it has not been built (no fallback)."""
    global _cu
    if _cu is None:
        L = C.CDLL(LIB_CU, mode=C.RTLD_GLOBAL)
        L.x265cu_open.argtypes = [C.POINTER(Config), C.POINTER(C.c_void_p)]
        L.x265cu_close.argtypes = [C.c_void_p]
        L.x265cu_last_error.restype = C.c_char_p
        L.x265cu_last_error.argtypes = [C.c_void_p]
        L.x265cu_get_geometry.argtypes = [C.c_void_p, C.POINTER(Geometry)]
        L.x265cu_sync.argtypes = [C.c_void_p]
        L.x265cu_host_register.argtypes = [C.c_void_p, C.c_size_t]
        L.x265cu_host_unregister.argtypes = [C.c_void_p]
        L.x265cu_frame_init.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_ssize_t, C.c_int, C.c_void_p]
        L.x265cu_frame_set_invqscale.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.x265cu_frame_var.argtypes = [C.c_void_p, C.c_void_p, C.c_ssize_t, C.c_void_p, C.c_void_p, C.c_ssize_t, C.c_int, C.c_void_p, C.c_void_p]
        L.x265cu_intra.argtypes = [C.c_void_p, C.c_int, C.POINTER(IntraOut)]
        L.x265cu_weight_cost_batch.argtypes = [C.c_void_p, C.c_int, C.POINTER(WeightItem), C.c_void_p]
        L.x265cu_estimate_batch.argtypes = [C.c_void_p, C.c_int, C.POINTER(Job), C.POINTER(JobResult)]
        L.x265cu_pixelcmp_batch.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_ssize_t, C.c_void_p, C.c_size_t, C.c_ssize_t,
                                            C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.x265cu_pixelcmp_frames.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_float)]
        L.x265cu_pixelcmp_planes.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_float)]
        L.x265cu_int_peak.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.x265cu_cutree_run.argtypes = [C.c_void_p, C.c_int, C.POINTER(CutreeOp), C.c_int, C.c_void_p, C.c_void_p]
        L.x265cu_frame_set_propagate.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.x265cu_frame_set_array.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.x265cu_stats_enable.argtypes = [C.c_void_p, C.c_int]
        L.x265cu_stats_get.argtypes = [C.c_void_p, C.POINTER(Stats), C.c_int]
        _cu = L
    return _cu


def lib_host():
    global _host
    if _host is None:
        lib_cu()
        L = C.CDLL(LIB_HOST)
        L.x265cuh_open.restype = C.c_void_p
        L.x265cuh_open.argtypes = [C.POINTER(HostParams), C.c_char_p, C.c_int]
        L.x265cuh_close.argtypes = [C.c_void_p]
        L.x265cuh_set_resident.argtypes = [C.c_void_p, C.c_int]
        L.x265cuh_sync.argtypes = [C.c_void_p]
        L.x265cuh_frame_slot.argtypes = [C.c_void_p]
        L.x265cuh_ctx.restype = C.c_void_p
        L.x265cuh_ctx.argtypes = [C.c_void_p]
        L.x265cuh_info.argtypes = [C.c_void_p, C.c_void_p]
        L.x265cuh_mvcost_crc.restype = C.c_uint32
        L.x265cuh_mvcost_crc.argtypes = [C.c_void_p]
        L.x265cuh_frame_alloc.restype = C.c_void_p
        L.x265cuh_frame_alloc.argtypes = [C.c_void_p]
        L.x265cuh_frame_free.argtypes = [C.c_void_p, C.c_void_p]
        L.x265cuh_pre_lookahead.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_ssize_t, C.c_void_p, C.c_void_p, C.c_ssize_t, C.c_int, C.c_int]
        L.x265cuh_pre_lookahead_batch.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 7 + [C.c_int]
        L.x265cuh_add_pictures.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 6
        L.x265cuh_estimate.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.x265cuh_array.restype = C.c_void_p
        L.x265cuh_array.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_size_t)]
        L.x265cuh_frame_scalars.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.x265cuh_cutree_zero.argtypes = [C.c_void_p, C.c_void_p]
        L.x265cuh_cutree_propagate.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double]
        L.x265cuh_cutree_finish.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_int]
        L.x265cuh_cutree_stats.argtypes = [C.c_void_p, C.c_void_p]
        L.x265cuh_cutree_sequence.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 8
        L.x265cuh_crc32.restype = C.c_uint32
        L.x265cuh_crc32.argtypes = [C.c_void_p, C.c_size_t]
        L.x265cuh_error.restype = C.c_char_p
        L.x265cuh_error.argtypes = [C.c_void_p]
        _host = L
    return _host


class Lookahead:
    """Python handle on x265cu::Lookahead (host/lookahead_cu.h)."""

    def __init__(self, width, height, depth=8, bframes=4, lookahead=20, slices=8, pool=16, weightp=1, aq_mode=1,
                 aq_strength=1.0, bframe_bias=0, device=0, slots=0, ctu=64, stream=None, search_warps=0,
                 fps_num=30, fps_denom=1, qcompress=0.6, weighted_bipred=0):
        self.L = lib_host()
        p = HostParams(width, height, depth, ctu, bframes, lookahead, slices, pool, weightp, aq_mode, aq_strength, bframe_bias, device, slots,
                       stream, search_warps, fps_num, fps_denom, qcompress, weighted_bipred)
        err = C.create_string_buffer(512)
        self.h = self.L.x265cuh_open(C.byref(p), err, 512)
        if not self.h:
            raise RuntimeError("x265cu::Lookahead::create failed: " + err.value.decode())
        info = (C.c_int * 16)()
        self.L.x265cuh_info(self.h, info)
        (self.wCU, self.hCU, self.nCU, self.stride, self.planeSize, self.numCoopSlices, self.numRowsPerSlice, self.lam,
         self.pixelBytes, self.paddedLines, self.padOffset, self.blocks) = list(info)[:12]
        self.depth = depth
        self.ctx = self.L.x265cuh_ctx(self.h)

    def close(self):
        if self.h:
            self.L.x265cuh_close(self.h)
            self.h = None

    def error(self):
        return self.L.x265cuh_error(self.h).decode()

    def mvcost_crc(self):
        return int(self.L.x265cuh_mvcost_crc(self.h))

    def frame_alloc(self):
        f = self.L.x265cuh_frame_alloc(self.h)
        if not f:
            raise RuntimeError("allocLowres failed: " + self.error())
        return f

    def frame_free(self, f):
        self.L.x265cuh_frame_free(self.h, f)

    def pre_lookahead(self, frame, y, u, v, poc, planes_back=True):
        r = self.L.x265cuh_pre_lookahead(self.h, frame, y.ctypes.data, y.strides[0] // y.itemsize,
                                         u.ctypes.data if u is not None else None, v.ctypes.data if v is not None else None,
                                         (u.strides[0] // u.itemsize) if u is not None else 0, poc, 1 if planes_back else 0)
        if r:
            raise RuntimeError("preLookahead failed: " + self.error())

    def sync(self):
        """wait for everything in flight, including the asynchronous plane copy-backs"""
        if self.L.x265cuh_sync(self.h):
            raise RuntimeError("sync failed: " + self.error())

    def set_resident(self, on):
        """inputs become device pointers, result arrays stay in HBM (only sums return)"""
        self.L.x265cuh_set_resident(self.h, 1 if on else 0)

    def prepare_pre_lookahead_batch(self, items):
        """pre-marshal a PreLookaheadGroup list: items = [(frame, y, ys, u, v, cs, poc)] with raw pointers"""
        n = len(items)
        return (n, (C.c_void_p * n)(*[i[0] for i in items]), (C.c_void_p * n)(*[i[1] for i in items]),
                (C.c_ssize_t * n)(*[i[2] for i in items]), (C.c_void_p * n)(*[i[3] for i in items]),
                (C.c_void_p * n)(*[i[4] for i in items]), (C.c_ssize_t * n)(*[i[5] for i in items]), (C.c_int * n)(*[i[6] for i in items]))

    def pre_lookahead_batch_prepared(self, prep, planes_back):
        n, fr, y, ys, u, v, cs, pocs = prep
        if self.L.x265cuh_pre_lookahead_batch(self.h, n, fr, y, ys, u, v, cs, pocs, 1 if planes_back else 0):
            raise RuntimeError("pre_lookahead_batch failed: " + self.error())

    def add_pictures_prepared(self, prep):
        """Lookahead::addPicture for the frames of a prepared list: their uploads start now (asynchronous)"""
        n, fr, y, ys, u, v, cs, _pocs = prep
        if self.L.x265cuh_add_pictures(self.h, n, fr, y, ys, u, v, cs):
            raise RuntimeError("addPicture failed: " + self.error())

    def pre_lookahead_ptr(self, frame, y, ys, u, v, cs, poc, planes_back):
        """raw-pointer form (host or, in resident mode, device addresses)"""
        r = self.L.x265cuh_pre_lookahead(self.h, frame, y, ys, u, v, cs, poc, 1 if planes_back else 0)
        if r:
            raise RuntimeError("preLookahead failed: " + self.error())

    def prepare_estimate(self, frames, triples):
        """pre-marshal an estimate call (bench inner loop): returns an opaque tuple for estimate_prepared"""
        n = len(triples)
        fr = (C.c_void_p * len(frames))(*frames)
        tr = (C.c_int * (3 * n))(*[x for t in triples for x in t])
        sc = (C.c_int64 * n)()
        return (fr, len(frames), tr, n, sc)

    def estimate_prepared(self, prep, batch):
        fr, nf, tr, n, sc = prep
        if self.L.x265cuh_estimate(self.h, fr, nf, tr, n, 1 if batch else 0, sc):
            raise RuntimeError("estimate failed: " + self.error())
        return sc

    def estimate(self, frames, triples, batch):
        """frames: list of frame handles; triples: list of (p0, p1, b) indices into frames."""
        n = len(triples)
        fr = (C.c_void_p * len(frames))(*frames)
        tr = (C.c_int * (3 * n))(*[x for t in triples for x in t])
        sc = (C.c_int64 * n)()
        r = self.L.x265cuh_estimate(self.h, fr, len(frames), tr, n, 1 if batch else 0, sc)
        if r:
            raise RuntimeError("estimate failed: " + self.error())
        return list(sc)

    def array(self, frame, which, d0=0, d1=0, dtype=np.uint8):
        nb = C.c_size_t()
        p = self.L.x265cuh_array(self.h, frame, which, d0, d1, C.byref(nb))
        if not p or not nb.value:
            return None
        buf = (C.c_char * nb.value).from_address(p)
        return np.frombuffer(buf, dtype=dtype)

    def crc(self, frame, which, d0=0, d1=0):
        nb = C.c_size_t()
        p = self.L.x265cuh_array(self.h, frame, which, d0, d1, C.byref(nb))
        if not p or not nb.value:
            return 0
        return int(self.L.x265cuh_crc32(p, nb.value))

    # ---- cuTree propagation (host/lookahead_cu.h: cuTreeZero / estimateCUPropagate / cuTreeFinish)
    def cutree_zero(self, frame):
        self.L.x265cuh_cutree_zero(self.h, frame)

    def cutree_propagate(self, frames, p0, p1, b, referenced, average_duration):
        fr = (C.c_void_p * len(frames))(*frames)
        if self.L.x265cuh_cutree_propagate(self.h, fr, len(frames), p0, p1, b, referenced, average_duration):
            raise RuntimeError("estimateCUPropagate failed: " + self.error())

    def cutree_finish(self, frame, average_duration, ref0_distance):
        if self.L.x265cuh_cutree_finish(self.h, frame, average_duration, ref0_distance):
            raise RuntimeError("cuTreeFinish failed: " + self.error())

    def prepare_cutree_sequence(self, calls):
        """pre-marshal a run of cuTree calls: ("M", frame) | ("X", fenc, ref0, ref1, d0, d1, referenced, avgDuration) |
        ("F", frame, avgDuration, ref0Distance)"""
        n = len(calls)
        kind, fenc, r0, r1 = (C.c_int * n)(), (C.c_void_p * n)(), (C.c_void_p * n)(), (C.c_void_p * n)()
        d0, d1, arg, dur = (C.c_int * n)(), (C.c_int * n)(), (C.c_int * n)(), (C.c_double * n)()
        for i, c in enumerate(calls):
            if c[0] == "M":
                kind[i], fenc[i] = 0, c[1]
            elif c[0] == "X":
                kind[i], fenc[i], r0[i], r1[i], d0[i], d1[i], arg[i], dur[i] = 1, c[1], c[2], c[3], c[4], c[5], c[6], c[7]
            else:
                kind[i], fenc[i], dur[i], arg[i] = 2, c[1], c[2], c[3]
        return (n, kind, fenc, r0, r1, d0, d1, arg, dur)

    def cutree_sequence_prepared(self, prep):
        if self.L.x265cuh_cutree_sequence(self.h, *prep):
            raise RuntimeError("cuTree sequence failed: " + self.error())

    def cutree_stats(self):
        o = (C.c_int64 * 3)()
        self.L.x265cuh_cutree_stats(self.h, o)
        return dict(zip(("steps", "runs", "reuploads"), list(o)))

    def scalars(self, frame, d0, d1):
        o = (C.c_int64 * 9)()
        self.L.x265cuh_frame_scalars(frame, d0, d1, o)
        keys = ("costEst", "costEstAq", "intraMbs", "wp_ssd", "wp_sum", "weighted", "wscale", "wdenom", "woffset")
        return dict(zip(keys, list(o)))

    def stats(self, reset=False):
        s = Stats()
        lib_cu().x265cu_stats_get(self.ctx, C.byref(s), 1 if reset else 0)
        return {"ms": dict(zip(K_NAMES, list(s.ms))), "launches": dict(zip(K_NAMES, list(s.launches))),
                "h2d": int(s.h2dBytes), "d2h": int(s.d2hBytes)}

    def stats_enable(self, timing=True):
        lib_cu().x265cu_stats_enable(self.ctx, 1 if timing else 0)
