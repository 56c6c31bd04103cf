"""Writes tests/golden/me_golden.json: results of the UNMODIFIED reference's MotionEstimate::motionEstimate
(oracle/_ref/libx265ref<depth>.so, built by build_ref.py from /root/reference) on the cases of harness/me_cases.py.
Run in the build container only; the JSON travels."""
import ctypes as C
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from harness import me_cases as mc   # noqa: E402
from oracle import pyoracle as po    # noqa: E402


def main():
    out = {"lut": "harness.me_cases.make_lut(20)", "depths": {}}
    lut = mc.make_lut(20)
    for depth in (8, 10):
        R = po.ref(depth)
        V, I, S = C.c_void_p, C.c_int, C.c_ssize_t
        R.x265ref_motion_estimate_batch.argtypes = [I, I, I, I, I, V, S, V, S, I, V, V, V]
        R.x265ref_motion_estimate_batch.restype = None
        res = []
        for case in mc.golden_cases():
            fenc, ref, stride, its = mc.build_case(case, depth)
            n = case["n"]
            want = (mc.MeResult * n)()
            R.x265ref_motion_estimate_batch(mc.METHODS[case["method"]], case["subme"], 0, case["w"], case["h"], fenc.ctypes.data, stride,
                                            ref.ctypes.data, stride, n, C.addressof(its), C.addressof(want), lut.ctypes.data + 65536 * 2)
            res.append(mc.results_list(want, n))
        out["depths"][str(depth)] = res
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "me_golden.json")
    with open(path, "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
