"""ctypes bindings for the CPU oracle (oracle/x265la_oracle.c), the reference shim
(oracle/_ref/libx265ref<depth>.so, only where it was built) and the trace format.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs.  Never imported by src/x265_b200.
"""
import ctypes as C
import os
import subprocess
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD = os.path.join(HERE, "_build")
REF = os.path.join(HERE, "_ref")
BFMAX = 16


class Geom(C.Structure):
    _fields_ = [("srcW", C.c_int), ("srcH", C.c_int), ("width", C.c_int), ("lines", C.c_int), ("stride", C.c_int),
                ("marginX", C.c_int), ("marginY", C.c_int), ("paddedLines", C.c_int),
                ("wCU", C.c_int), ("hCU", C.c_int), ("nCU", C.c_int),
                ("planeSize", C.c_int64), ("padOffset", C.c_int64)]


class Weight(C.Structure):
    _fields_ = [("present", C.c_int), ("scale", C.c_int), ("denom", C.c_int), ("offset", C.c_int)]


class Frame(C.Structure):
    _fields_ = [("g", Geom), ("bframes", C.c_int), ("frameNum", C.c_int), ("hasAq", C.c_int),
                ("buffer", C.c_void_p * 4), ("plane", C.c_void_p * 4),
                ("intraCost", C.c_void_p), ("intraMode", C.c_void_p), ("invQscale", C.c_void_p),
                ("qpAqOffset", C.c_void_p), ("qpCuTreeOffset", C.c_void_p), ("blockVariance", C.c_void_p),
                ("lowresCosts", (C.c_void_p * (BFMAX + 2)) * (BFMAX + 2)),
                ("rowSatds", (C.c_void_p * (BFMAX + 2)) * (BFMAX + 2)),
                ("mvs", (C.c_void_p * (BFMAX + 1)) * 2),
                ("mvCosts", (C.c_void_p * (BFMAX + 1)) * 2),
                ("costEst", (C.c_int64 * (BFMAX + 2)) * (BFMAX + 2)),
                ("costEstAq", (C.c_int64 * (BFMAX + 2)) * (BFMAX + 2)),
                ("intraMbs", C.c_int * (BFMAX + 2)),
                ("wp_ssd", C.c_uint64 * 3), ("wp_sum", C.c_uint64 * 3), ("frameVariance", C.c_uint64),
                ("weightedCostDelta", C.c_double * (BFMAX + 2)),
                ("propagateCost", C.c_void_p)]


def build_oracle():
    """(re)build liboracle8/10.so with the Makefile; cheap when up to date."""
    subprocess.run(["make", "-s", "-C", HERE], check=True)


_libs = {}


def oracle(depth=8, emul=False):
    """Load liboracle<depth>.so (building it if needed).  emul=True loads libcoreemul<depth>.so
    instead: the same oracle plus emul_estimate(), the CPU emulation of the product's la_core.h
    state machine (tests/core_emul.cpp)."""
    key = ("e" if emul else "o", depth)
    if key in _libs:
        return _libs[key]
    path = os.path.join(BUILD, ("libcoreemul%d.so" if emul else "liboracle%d.so") % depth)
    if not os.path.exists(path):
        build_oracle()
    L = C.CDLL(path)
    if emul:
        L.emul_estimate.restype = C.c_int64
        L.emul_estimate.argtypes = [C.c_void_p, C.POINTER(Frame), C.POINTER(Frame), C.POINTER(Frame)] + [C.c_int] * 6 + \
                                   [C.POINTER(Weight), C.POINTER(Weight)]
    L.ola_frame_create.restype = C.POINTER(Frame)
    L.ola_frame_create.argtypes = [C.c_int] * 6
    L.ola_frame_destroy.argtypes = [C.POINTER(Frame)]
    L.ola_ctx_create.restype = C.c_void_p
    L.ola_ctx_create.argtypes = [C.c_int] * 3
    L.ola_ctx_destroy.argtypes = [C.c_void_p]
    L.ola_frame_init.argtypes = [C.POINTER(Frame), C.c_void_p, C.c_ssize_t, C.c_int]
    L.ola_aq_frame.argtypes = [C.POINTER(Frame), C.c_void_p, C.c_ssize_t, C.c_void_p, C.c_void_p, C.c_ssize_t,
                               C.c_int, C.c_double, C.c_int]
    L.ola_intra_estimate.argtypes = [C.POINTER(Frame), C.c_int]
    L.ola_estimate.restype = C.c_int64
    L.ola_estimate.argtypes = [C.c_void_p, C.POINTER(Frame), C.POINTER(Frame), C.POINTER(Frame)] + [C.c_int] * 6 + \
                              [C.POINTER(Weight), C.POINTER(Weight)]
    L.ola_propagate_cost.argtypes = [C.c_void_p] * 5 + [C.POINTER(C.c_double), C.c_int]
    L.ola_cutree_zero.argtypes = [C.POINTER(Frame)]
    L.ola_estimate_cu_propagate.argtypes = [C.POINTER(Frame)] * 3 + [C.c_int] * 3 + [C.c_double, C.c_int, C.c_int, C.c_int]
    L.ola_cutree_finish.argtypes = [C.POINTER(Frame), C.c_double, C.c_int, C.c_int, C.c_int, C.c_double]
    L.ola_weights_analyse.argtypes = [C.c_void_p, C.POINTER(Frame), C.POINTER(Frame), C.POINTER(Weight)]
    L.ola_weight_cost_luma.restype = C.c_uint32
    L.ola_weight_cost_luma.argtypes = [C.c_void_p, C.POINTER(Frame), C.POINTER(Frame), C.POINTER(Weight)]
    for n in ("ola_sad8x8", "ola_satd8x8", "ola_satd4x4", "ola_sa8d8x8", "ola_sa8d16x16"):
        getattr(L, n).argtypes = [C.c_void_p, C.c_ssize_t, C.c_void_p, C.c_ssize_t]
    for n in ("ola_pu_sad", "ola_pu_satd"):
        getattr(L, n).argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_ssize_t, C.c_void_p, C.c_ssize_t]
    L.ola_pixelavg8x8.argtypes = [C.c_void_p, C.c_ssize_t, C.c_void_p, C.c_ssize_t, C.c_void_p, C.c_ssize_t]
    L.ola_frame_init_lowres.argtypes = [C.c_void_p] * 5 + [C.c_ssize_t, C.c_ssize_t, C.c_int, C.c_int]
    L.ola_extend_border.argtypes = [C.c_void_p, C.c_ssize_t] + [C.c_int] * 4
    L.ola_intra_filter8.argtypes = [C.c_void_p, C.c_void_p]
    L.ola_intra_pred8.argtypes = [C.c_int, C.c_void_p, C.c_ssize_t, C.c_void_p, C.c_int]
    L.ola_weight_pp.argtypes = [C.c_void_p, C.c_void_p, C.c_ssize_t] + [C.c_int] * 6
    L.ola_var16.restype = C.c_uint64
    L.ola_var16.argtypes = [C.c_void_p, C.c_ssize_t]
    L.ola_var8.restype = C.c_uint64
    L.ola_var8.argtypes = [C.c_void_p, C.c_ssize_t]
    L.ola_exp2fix8.argtypes = [C.c_double]
    L.ola_mvcost_table.argtypes = [C.c_void_p]
    L.ola_crc32.restype = C.c_uint32
    L.ola_crc32.argtypes = [C.c_void_p, C.c_size_t]
    L.ola_synth_frame.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
    L.ola_copy_picture.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_ssize_t]
    L.ola_geometry.argtypes = [C.c_int] * 4 + [C.POINTER(Geom)]
    L.ola_coop_slices.argtypes = [C.c_int] * 4 + [C.POINTER(C.c_int), C.POINTER(C.c_int)]
    _libs[key] = L
    return L


def ref_available(depth=8):
    return os.path.exists(os.path.join(REF, "libx265ref%d.so" % depth))


def ref(depth=8):
    """Load the reference shim (oracle/_ref); only exists where build_ref.py ran."""
    key = ("r", depth)
    if key in _libs:
        return _libs[key]
    L = C.CDLL(os.path.join(REF, "libx265ref%d.so" % depth))
    L.x265ref_setup()
    for n in ("x265ref_sad8x8", "x265ref_satd8x8", "x265ref_sa8d8x8", "x265ref_sa8d16x16"):
        getattr(L, n).argtypes = [C.c_void_p, C.c_ssize_t, C.c_void_p, C.c_ssize_t]
    for n in ("x265ref_pu_sad", "x265ref_pu_satd"):
        getattr(L, n).argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_ssize_t, C.c_void_p, C.c_ssize_t]
    L.x265ref_sad_x3_8x8.argtypes = [C.c_void_p] * 4 + [C.c_ssize_t, C.c_void_p]
    L.x265ref_sad_x4_8x8.argtypes = [C.c_void_p] * 5 + [C.c_ssize_t, C.c_void_p]
    L.x265ref_pixelavg8x8.argtypes = [C.c_void_p, C.c_ssize_t, C.c_void_p, C.c_ssize_t, C.c_void_p, C.c_ssize_t]
    L.x265ref_frame_init_lowres.argtypes = [C.c_void_p] * 5 + [C.c_ssize_t, C.c_ssize_t, C.c_int, C.c_int]
    L.x265ref_extend_border.argtypes = [C.c_void_p, C.c_ssize_t] + [C.c_int] * 4
    L.x265ref_intra_filter8.argtypes = [C.c_void_p, C.c_void_p]
    L.x265ref_intra_pred8.argtypes = [C.c_int, C.c_void_p, C.c_ssize_t, C.c_void_p, C.c_int]
    L.x265ref_weight_pp.argtypes = [C.c_void_p, C.c_void_p, C.c_ssize_t] + [C.c_int] * 6
    L.x265ref_var16.restype = C.c_uint64
    L.x265ref_var16.argtypes = [C.c_void_p, C.c_ssize_t]
    L.x265ref_var8.restype = C.c_uint64
    L.x265ref_var8.argtypes = [C.c_void_p, C.c_ssize_t]
    L.x265ref_exp2fix8.argtypes = [C.c_double]
    L.x265ref_propagate_cost.argtypes = [C.c_void_p] * 5 + [C.POINTER(C.c_double), C.c_int]
    L.x265ref_mvcost_table.argtypes = [C.c_void_p]
    L.x265ref_run_lookahead.restype = C.c_double
    L.x265ref_run_lookahead.argtypes = [C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                        C.c_char_p, C.c_char_p, C.c_void_p, C.c_void_p]
    _libs[key] = L
    return L


def ref_run_lookahead(depth, width, height, nframes, seed, opts, pool_threads, trace=None, dump=None):
    """Run the reference's own Lookahead on the synthetic clip.  opts: list of (name, value).
    Returns (seconds, slice_types[nframes], stats dict)."""
    L = ref(depth)
    names = (C.c_char_p * len(opts))(*[k.encode() for k, _ in opts])
    vals = (C.c_char_p * len(opts))(*[(str(v).encode() if v is not None else None) for _, v in opts])
    st = (C.c_int * nframes)()
    stats = (C.c_long * 8)()
    secs = L.x265ref_run_lookahead(width, height, nframes, seed, names, vals, len(opts), pool_threads,
                                   trace.encode() if trace else None, dump.encode() if dump else None, st, stats)
    if secs < 0:
        raise RuntimeError("x265ref_run_lookahead failed: %r" % secs)
    keys = ("pre", "jobs", "searchL0", "searchL1", "batches", "decided", "propagates")
    return secs, list(st), dict(zip(keys, list(stats)[:7]))


# ------------------------------------------------------------------------------------------
# numpy views on oracle frames
# ------------------------------------------------------------------------------------------
def pixel_dtype(depth):
    return np.uint8 if depth == 8 else np.uint16


def _view(ptr, dtype, n):
    if not ptr:
        return None
    buf = (C.c_char * (n * np.dtype(dtype).itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype, count=n)


class OFrame:
    """numpy-friendly wrapper of an oracle ola_frame*."""

    def __init__(self, lib, depth, srcW, srcH, marginX, marginY, bframes, aq):
        self.lib, self.depth = lib, depth
        self.p = lib.ola_frame_create(srcW, srcH, marginX, marginY, bframes, 1 if aq else 0)
        self.c = self.p.contents
        self.g = self.c.g
        self.dt = pixel_dtype(depth)

    def close(self):
        if self.p:
            self.lib.ola_frame_destroy(self.p)
            self.p = None

    def planes(self):
        return _view(self.c.buffer[0], self.dt, 4 * self.g.planeSize)

    def intra_cost(self):
        return _view(self.c.intraCost, np.int32, self.g.nCU)

    def intra_mode(self):
        return _view(self.c.intraMode, np.uint8, self.g.nCU)

    def inv_qscale(self):
        return _view(self.c.invQscale, np.int32, self.g.nCU)

    def lowres_costs(self, d0, d1):
        return _view(self.c.lowresCosts[d0][d1], np.uint16, self.g.nCU)

    def row_satds(self, d0, d1):
        return _view(self.c.rowSatds[d0][d1], np.int32, self.g.hCU)

    def mvs(self, lst, d):
        return _view(self.c.mvs[lst][d - 1], np.int16, 2 * self.g.nCU)

    def mv_costs(self, lst, d):
        return _view(self.c.mvCosts[lst][d - 1], np.int32, self.g.nCU)

    def propagate_cost(self):
        return _view(self.c.propagateCost, np.uint16, self.g.nCU)

    def qp_cutree_offset(self):
        return _view(self.c.qpCuTreeOffset, np.float64, self.g.nCU)


def crc(a):
    return zlib.crc32(np.ascontiguousarray(a).tobytes()) & 0xFFFFFFFF


def synth_padded(lib, depth, w, h, t, nframes, seed):
    """Synthetic frame t padded like PicYuv::copyFromPicture.  Returns (Y, U, V) 2-D arrays whose
    [0:h, 0:w] (chroma: half) region is the picture and whose extra right/bottom samples are the
    replicated edge."""
    dt = pixel_dtype(depth)
    y = np.zeros((h, w), dt)
    u = np.zeros((h // 2, w // 2), dt)
    v = np.zeros((h // 2, w // 2), dt)
    lib.ola_synth_frame(w, h, t, nframes, seed, y.ctypes.data, w, u.ctypes.data, v.ctypes.data, w // 2)
    return pad_picture(y, 1), pad_picture(u, 0, chroma_of=(w, h)), pad_picture(v, 0, chroma_of=(w, h))


def pad_picture(p, extra, chroma_of=None):
    """PicYuv::copyFromPicture edge padding (picyuv.cpp:168-178,287-335): to a multiple of 16 luma
    samples, plus one luma column/row; chroma gets (padx >> 1, pady >> 1)."""
    h, w = p.shape
    if chroma_of is None:
        padx = (16 - (w & 15)) & 15
        pady = (16 - (h & 15)) & 15
        padx += 1
        pady += 1
    else:
        lw, lh = chroma_of
        padx = (((16 - (lw & 15)) & 15) + 1) >> 1
        pady = (((16 - (lh & 15)) & 15) + 1) >> 1
    # rows live in a wider backing array, like PicYuv's planes inside their margins (picyuv.cpp:62-75): the view's
    # row pitch is a multiple of 64 samples, as an encoder's picture buffers are
    margin = 96 if chroma_of is None else 48
    pitch = (w + padx + 2 * margin + 63) // 64 * 64
    backing = np.zeros((h + pady, pitch), p.dtype)
    out = backing[:, margin:margin + w + padx]
    out[:h, :w] = p
    out[:h, w:] = p[:, w - 1:w]
    out[h:, :] = out[h - 1:h, :]
    return out


# ------------------------------------------------------------------------------------------
# trace files (written by oracle/ref_shim.cpp, committed under tests/golden/)
# ------------------------------------------------------------------------------------------
class Trace:
    def __init__(self, path):
        # ("P", dict) | ("B", [jobs]) | ("J", dict) | ("D", poc, type) | cuTree: ("M", poc) | ("X", dict) | ("F", dict)
        self.events = []
        self.cfg = {}
        batch = None
        with open(path) as f:
            for line in f:
                t = line.split()
                if not t or t[0].startswith("#"):
                    continue
                k = t[0]
                if k == "C":
                    names = ("width", "height", "depth", "nframes", "seed", "bframes", "lookahead", "badapt", "weightp",
                             "aqmode", "cutree", "numCoopSlices", "numRowsPerSlice", "bFrameBias", "pool")
                    self.cfg.update(dict(zip(names, [int(x) for x in t[1:]])))
                elif k == "Q":
                    self.cfg["aqStrength"] = float(t[1])
                    self.cfg["scenecut"] = int(t[2])
                    self.cfg["keyint"] = int(t[3])
                elif k == "L":
                    self.cfg["lutCrc"] = int(t[1], 16)
                    self.cfg["lambda"] = int(t[2])
                elif k == "T":
                    self.cfg.update(fpsNum=int(t[1]), fpsDenom=int(t[2]), qCompress=float(t[3]), weightedBiPred=int(t[4]),
                                    vbvBufferSize=int(t[5]))
                elif k == "M":
                    self.events.append(("M", int(t[1])))
                elif k == "X":
                    self.events.append(("X", dict(p0=int(t[1]), b=int(t[2]), p1=int(t[3]), referenced=int(t[4]),
                                                   avgDuration=float(t[5]), ref0=int(t[6], 16), ref1=int(t[7], 16), own=int(t[8], 16))))
                elif k == "F":
                    self.events.append(("F", dict(poc=int(t[1]), ref0Distance=int(t[2]), avgDuration=float(t[3]),
                                                   qpCuTreeOffset=int(t[4], 16), propagateCost=int(t[5], 16))))
                elif k == "P":
                    self.events.append(("P", dict(poc=int(t[1]), planes=int(t[2], 16), intraCost=int(t[3], 16),
                                                   intraMode=int(t[4], 16), lowresCosts=int(t[5], 16), rowSatds=int(t[6], 16),
                                                   costEst=int(t[7]), costEstAq=int(t[8]), invQ=int(t[9], 16),
                                                   wp_ssd=int(t[10]), wp_sum=int(t[11]))))
                elif k == "B":
                    batch = []
                elif k == "E":
                    self.events.append(("B", batch))
                    batch = None
                elif k == "J":
                    j = dict(p0=int(t[1]), b=int(t[2]), p1=int(t[3]), s0=int(t[4]), s1=int(t[5]), batch=int(t[6]),
                             sliced=int(t[7]), wflag=int(t[8]), wscale=int(t[9]), wdenom=int(t[10]), woffset=int(t[11]),
                             costEst=int(t[12]), costEstAq=int(t[13]), intraMbs=int(t[14]),
                             mv0=int(t[15], 16), mv1=int(t[16], 16), mc0=int(t[17], 16), mc1=int(t[18], 16),
                             lowresCosts=int(t[19], 16), rowSatds=int(t[20], 16), wdelta=float(t[21]))
                    if batch is not None:
                        batch.append(j)
                    else:
                        self.events.append(("J", j))
                elif k == "D":
                    self.events.append(("D", int(t[1]), int(t[2])))

    def jobs(self):
        for e in self.events:
            if e[0] == "J":
                yield e[1]
            elif e[0] == "B":
                for j in e[1]:
                    yield j


def read_dump(path):
    """Binary dump of the reference's arrays -> dict[(tag, a, b, c, id)] = bytes."""
    out = {}
    with open(path, "rb") as f:
        data = f.read()
    off = 0
    while off < len(data):
        tag = data[off:off + 4].decode()
        a, b, c, idx, n = np.frombuffer(data, np.int32, 5, off + 4)
        off += 24
        out[(tag, int(a), int(b), int(c), int(idx))] = data[off:off + int(n)]
        off += int(n)
    return out


class OracleReplay:
    """Replays a reference trace through the oracle and compares every checksum.

    This is the pin of the oracle at clip level: the reference produced the trace (and the
    checksums in it) with its own compiled C code; the oracle must reproduce all of them from the
    same synthetic input."""

    def __init__(self, trace, use_trace_weights=False, keep=None, emul=False):
        self.t = trace
        cfg = trace.cfg
        self.depth = cfg["depth"]
        self.lib = oracle(self.depth, emul=emul)
        self.estimate = self.lib.emul_estimate if emul else self.lib.ola_estimate
        self.ctx = self.lib.ola_ctx_create(cfg["bFrameBias"], cfg["numCoopSlices"], cfg["numRowsPerSlice"])
        self.frames = {}
        self.use_trace_weights = use_trace_weights
        self.keep = keep if keep is not None else cfg["lookahead"] + cfg["bframes"] + 8
        self.mismatches = []
        self.njobs = 0
        self.npropagate = 0

    def close(self):
        for f in self.frames.values():
            f.close()
        self.frames = {}
        self.lib.ola_ctx_destroy(self.ctx)

    def _check(self, what, got, want):
        if got != want:
            self.mismatches.append((what, got, want))

    def pre(self, e):
        cfg = self.t.cfg
        w, h = cfg["width"], cfg["height"]
        poc = e["poc"]
        f = OFrame(self.lib, self.depth, w, h, 96, 80, cfg["bframes"], cfg["aqmode"] != 0)
        y, u, v = synth_padded(self.lib, self.depth, w, h, poc, cfg["nframes"], cfg["seed"])
        self.lib.ola_frame_init(f.p, y.ctypes.data, y.strides[0] // y.itemsize, poc)
        if cfg["aqmode"] or cfg["weightp"]:
            self.lib.ola_aq_frame(f.p, y.ctypes.data, y.strides[0] // y.itemsize, u.ctypes.data, v.ctypes.data, u.strides[0] // u.itemsize,
                                  cfg["aqmode"], cfg["aqStrength"], cfg["weightp"])
        self.lib.ola_intra_estimate(f.p, 1 if self.depth == 8 else 16)
        tag = "P%d." % poc
        self._check(tag + "planes", crc(f.planes()), e["planes"])
        if cfg["aqmode"]:
            self._check(tag + "invQ", crc(f.inv_qscale()), e["invQ"])
        self._check(tag + "wp_ssd", int(f.c.wp_ssd[0]), e["wp_ssd"])
        self._check(tag + "wp_sum", int(f.c.wp_sum[0]), e["wp_sum"])
        self._check(tag + "intraCost", crc(f.intra_cost()), e["intraCost"])
        self._check(tag + "intraMode", crc(f.intra_mode()), e["intraMode"])
        self._check(tag + "lowresCosts", crc(f.lowres_costs(0, 0)), e["lowresCosts"])
        self._check(tag + "rowSatds", crc(f.row_satds(0, 0)), e["rowSatds"])
        self._check(tag + "costEst", int(f.c.costEst[0][0]), e["costEst"])
        self._check(tag + "costEstAq", int(f.c.costEstAq[0][0]), e["costEstAq"])
        self.frames[poc] = f
        for old in [p for p in self.frames if p < poc - self.keep]:
            self.frames.pop(old).close()

    def job(self, j):
        cfg = self.t.cfg
        fenc, r0, r1 = self.frames[j["b"]], self.frames[j["p0"]], self.frames[j["p1"]]
        d0, d1 = j["b"] - j["p0"], j["p1"] - j["b"]
        wt = None
        used = Weight()
        if self.use_trace_weights:
            wt = C.pointer(Weight(1 if j["wflag"] == 1 else 0, j["wscale"], j["wdenom"], j["woffset"]))
        score = self.estimate(self.ctx, fenc.p, r0.p, r1.p, d0, d1, -1, -1, 0 if j["batch"] else 1,
                                      cfg["weightp"], wt, C.byref(used))
        tag = "J%d/%d/%d." % (j["p0"], j["b"], j["p1"])
        want = j["costEst"] * 100 // (130 + cfg["bFrameBias"]) if d1 > 0 else j["costEst"]
        self._check(tag + "weight", (used.present, used.scale, used.denom, used.offset) if used.present else (0, 0, 0, 0),
                    (1, j["wscale"], j["wdenom"], j["woffset"]) if j["wflag"] == 1 else (0, 0, 0, 0))
        self._check(tag + "costEst", int(score), want)
        self._check(tag + "costEstAq", int(fenc.c.costEstAq[d0][d1]), j["costEstAq"])
        if d1 == 0:   # B-type jobs only log whatever a concurrent P-type job had accumulated so far
            self._check(tag + "intraMbs", int(fenc.c.intraMbs[d0]), j["intraMbs"])
        if d0 > 0:
            self._check(tag + "mv0", crc(fenc.mvs(0, d0)), j["mv0"])
            self._check(tag + "mc0", crc(fenc.mv_costs(0, d0)), j["mc0"])
        if d1 > 0:
            self._check(tag + "mv1", crc(fenc.mvs(1, d1)), j["mv1"])
            self._check(tag + "mc1", crc(fenc.mv_costs(1, d1)), j["mc1"])
        self._check(tag + "lowresCosts", crc(fenc.lowres_costs(d0, d1)), j["lowresCosts"])
        self._check(tag + "rowSatds", crc(fenc.row_satds(d0, d1)), j["rowSatds"])
        self.njobs += 1

    # ---- cuTree
    def ct_zero(self, poc):
        self.lib.ola_cutree_zero(self.frames[poc].p)

    def ct_propagate(self, x):
        cfg = self.t.cfg
        fenc, r0, r1 = self.frames[x["b"]], self.frames[x["p0"]], self.frames[x["p1"]]
        d0, d1 = x["b"] - x["p0"], x["p1"] - x["b"]
        self.lib.ola_estimate_cu_propagate(fenc.p, r0.p, r1.p, d0, d1, x["referenced"], x["avgDuration"], cfg["fpsNum"], cfg["fpsDenom"],
                                           cfg["weightedBiPred"])
        tag = "X%d/%d/%d." % (x["p0"], x["b"], x["p1"])
        self._check(tag + "ref0", crc(r0.propagate_cost()), x["ref0"])
        if d1 > 0:
            self._check(tag + "ref1", crc(r1.propagate_cost()), x["ref1"])
        own = fenc.propagate_cost()
        self._check(tag + "own", crc(own if x["referenced"] else own[:fenc.g.wCU]), x["own"])
        self.npropagate += 1

    def ct_finish(self, e):
        cfg = self.t.cfg
        f = self.frames[e["poc"]]
        self.lib.ola_cutree_finish(f.p, e["avgDuration"], cfg["fpsNum"], cfg["fpsDenom"], e["ref0Distance"], 5.0 * (1.0 - cfg["qCompress"]))
        self._check("F%d.propagateCost" % e["poc"], crc(f.propagate_cost()), e["propagateCost"])
        self._check("F%d.qpCuTreeOffset" % e["poc"], crc(f.qp_cutree_offset()), e["qpCuTreeOffset"])

    def run(self, max_events=None):
        n = 0
        for e in self.t.events:
            if e[0] == "P":
                self.pre(e[1])
            elif e[0] == "M":
                self.ct_zero(e[1])
            elif e[0] == "X":
                self.ct_propagate(e[1])
            elif e[0] == "F":
                self.ct_finish(e[1])
            elif e[0] == "J":
                self.job(e[1])
            elif e[0] == "B":
                for j in e[1]:
                    self.job(j)
            n += 1
            if max_events and n >= max_events:
                break
        return self.mismatches
