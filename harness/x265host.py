"""x265 1.9's own Lookahead driven frame by frame (no frame encoders) by harness/x265_la_driver.cpp, as either arm:

  gpu=False  oracle/_ref/libx265ref<depth>.so   the UNMODIFIED reference on the host cores (C primitives)
  gpu=True   oracle/_ref/libx265gpu<depth>.so   the same x265 objects with slicetype.cpp / lowres.cpp / picyuv.cpp bound to
                                                libx265cu.so (integration/): x265 is the host, the GPU does the estimates

Same driver, same synthetic clip, same pool size, same trace format; `compare_traces` is the parity check between the two.
"""
import ctypes as C
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")
_libs = {}


def lib_path(depth, gpu):
    return os.path.join(REF, "libx265%s%d.so" % ("gpu" if gpu else "ref", depth))


def available(depth, gpu):
    return os.path.exists(lib_path(depth, gpu))


def lib(depth, gpu):
    key = (depth, gpu)
    if key not in _libs:
        L = C.CDLL(lib_path(depth, gpu))
        L.x265la_open.restype = C.c_void_p
        L.x265la_open.argtypes = [C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int)]
        L.x265la_run.restype = C.c_double
        L.x265la_run.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_int, C.c_void_p, C.c_void_p]
        L.x265la_close.argtypes = [C.c_void_p]
        _libs[key] = L
    return _libs[key]


class LaDriver:
    """open: params, input Frames (synthetic clip in PicYuv); run(): one pass of a fresh Lookahead over the clip."""

    def __init__(self, depth, width, height, nframes, seed, opts, pool_threads, gpu):
        self.L = lib(depth, gpu)
        self.n = nframes
        names = (C.c_char_p * len(opts))(*[k.encode() for k, _ in opts])
        vals = (C.c_char_p * len(opts))(*[(str(v).encode() if v is not None else None) for _, v in opts])
        err = C.c_int(0)
        self.h = self.L.x265la_open(width, height, nframes, seed, names, vals, len(opts), pool_threads, C.byref(err))
        if not self.h:
            raise RuntimeError("x265la_open failed: %d" % err.value)

    def run(self, trace=None, dump=None, level=2):
        st = (C.c_int * self.n)()
        stats = (C.c_long * 8)()
        secs = self.L.x265la_run(self.h, trace.encode() if trace else None, dump.encode() if dump else None, level, st, stats)
        if secs < 0:
            raise RuntimeError("x265la_run failed: %r" % secs)
        keys = ("pre", "jobs", "searchL0", "searchL1", "batches", "decided", "propagates")
        return secs, list(st), dict(zip(keys, list(stats)[:7]))

    def close(self):
        if self.h:
            self.L.x265la_close(self.h)
            self.h = None


def _normalise(path, drop_x):
    """A trace as a sorted list of lines.  Estimates of one batch finish in any order on the reference's worker threads, so
    only the multiset of lines is defined; intraMbs[d0] of a B estimate (p1 != b) is whatever the P estimate of the same
    distance had added when the hook ran (thread timing in the reference itself) and is blanked."""
    out = []
    for ln in open(path).read().split("\n"):
        if not ln or ln[0] == "#":
            continue
        f = ln.split(" ")
        if f[0] == "X" and drop_x:
            continue
        if f[0] == "J" and f[2] != f[3]:
            f[14] = "-"
        out.append(" ".join(f))
    return sorted(out)


def compare_traces(path, golden, level=2):
    """-> list of (line only in `path`, line only in `golden`) differences; [] = every estimate, array CRC, cuTree step
    and slice-type decision equal"""
    a, b = _normalise(path, level < 2), _normalise(golden, level < 2)
    if a == b:
        return []
    sa, sb = set(a), set(b)
    return [("+", x) for x in a if x not in sb][:20] + [("-", x) for x in b if x not in sa][:20] or [("multiplicity", "", "")]
