"""x265 1.9's OWN lookahead as the host of the GPU path (oracle/_ref/libx265gpu<depth>.so = the reference's objects with
slicetype.cpp / lowres.cpp / picyuv.cpp / weightPrediction.cpp bound to libx265cu.so by integration/): the trace it writes --
every pre-lookahead frame, every estimate with the CRCs of all its output arrays, every cuTree step and cuTreeFinish, every
slice-type decision -- must equal the trace of the unmodified reference (tests/golden/*.trace), with the cuTree steps queued
(level 1: as in production) and fetched after every step (level 2: the per-step propagateCost CRCs are compared too)."""
import ctypes as C
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _have(depth):
    from harness import x265host as xh
    return xh.available(depth, True)


def test_gpu_host_library_exports_driver_and_binding():
    """no GPU needed: the library loads and carries the driver, the binding call-outs and the pinned allocator"""
    if not _have(8):
        pytest.skip("oracle/_ref/libx265gpu8.so not built (reference tree absent at build time)")
    from harness import x265host as xh
    L = xh.lib(8, True)
    for name in ("x265la_open", "x265la_run", "x265la_close", "x265glue_open", "x265glue_pre_list", "x265glue_ensure", "x265glue_finish_batch",
                 "x265glue_propagate", "x265glue_ct_fetch", "x265glue_wp_prepare", "x265glue_wp_cost", "x265glue_malloc", "x265glue_totals"):
        assert hasattr(L, name), name
    # the allocator works without a device (registration is only an optimisation)
    L.x265glue_malloc.restype = C.c_void_p
    L.x265glue_malloc.argtypes = [C.c_size_t]
    L.x265glue_free.argtypes = [C.c_void_p]
    p = L.x265glue_malloc(100000)
    assert p and p % 4096 == 0
    L.x265glue_free(p)


@pytest.mark.gpu
@pytest.mark.parametrize("name,levels", [("tiny8", (2, 1)), ("odd8", (2, 1)), ("tiny10", (2, 1)), ("c0_720p", (2, 1)), ("pool3_720p", (2, 1)),
                                         ("c0_720p10", (1,)), ("c0_1080p", (1,)), ("c1_1080p", (2, 1)), ("c2_4k", (1,)), ("c3_4k10", (1,))])
def test_x265_host_trace_equals_reference(name, levels, built, tmp_path):
    from harness import x265host as xh
    from harness.workloads import WORKLOADS
    depth, w, h, n, seed, pool, opts, _ = WORKLOADS[name]
    if not _have(depth):
        pytest.skip("oracle/_ref/libx265gpu%d.so not built (reference tree absent at build time)" % depth)
    d = xh.LaDriver(depth, w, h, n, seed, opts, pool, True)
    try:
        for level in levels:
            tr = str(tmp_path / ("%s_l%d.trace" % (name, level)))
            _, types, stats = d.run(trace=tr, level=level)
            assert stats["decided"] == n
            mm = xh.compare_traces(tr, os.path.join(ROOT, "tests", "golden", name + ".trace"), level)
            assert not mm, (name, level, mm[:6])
        # and once more without any observation: the production path must still decide every picture
        _, types2, _ = d.run()
        assert types2 == types
    finally:
        d.close()


@pytest.mark.gpu
@pytest.mark.parametrize("env", [{"X265CU_TRELLIS_AHEAD": "0"}, {"X265CU_GLUE_OWN_FINISH": "0"}, {"X265CU_LOOKAHEAD_CACHE": "0"}, {"X265CU_HOST_THREADS": "1"}])
def test_x265_host_switches_do_not_change_the_trace(env, built, tmp_path):
    """the other setting of each host-side switch of the x265-hosted path: the look-ahead cache without its rule-based part /
    off altogether, cuTreeFinish's mapping by x265's own loop, the AQ mapping on one thread -- same trace.  (A subprocess per
    setting: the switches are read once per process.)"""
    import subprocess
    import sys
    if not _have(8):
        pytest.skip("oracle/_ref/libx265gpu8.so not built (reference tree absent at build time)")
    code = (
        "import os, sys\n"
        "sys.path.insert(0, %r)\n"
        "from harness import x265host as xh\n"
        "from harness.workloads import WORKLOADS\n"
        "for name in ('c0_720p', 'pool3_720p'):\n"
        "    depth, w, h, n, seed, pool, opts, _ = WORKLOADS[name]\n"
        "    d = xh.LaDriver(depth, w, h, n, seed, opts, pool, True)\n"
        "    tr = os.path.join(%r, name + '.trace')\n"
        "    d.run(trace=tr, level=1)\n"
        "    mm = xh.compare_traces(tr, os.path.join(%r, 'tests', 'golden', name + '.trace'), 1)\n"
        "    d.close()\n"
        "    assert not mm, (name, mm[:6])\n"
        "print('ok')\n" % (ROOT, str(tmp_path), ROOT))
    e = dict(os.environ)
    e.update(env)
    r = subprocess.run([sys.executable, "-c", code], env=e, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert r.returncode == 0 and "ok" in r.stdout, (env, r.stdout[-400:], r.stderr[-800:])
