"""Time the reference's own CPU lookahead on a named workload (child process of bench.py).

usage: refrun.py <workload> <pool threads>   ->  one JSON line {"seconds", "threads", "kind", "frames"}

kind "reference": oracle/_ref/libx265ref<depth>.so, i.e. the UNMODIFIED x265 1.9 Lookahead class driven
frame by frame (no frame encoders), compiled from /root/reference by oracle/build_ref.py.
kind "port": only if that library is missing -- the single-threaded oracle replaying the trace.
The process leaves with os._exit: the reference's pool threads/static destructors do not survive
interpreter teardown reliably.
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from harness.workloads import WORKLOADS  # noqa: E402
from oracle import pyoracle as po  # noqa: E402


def main():
    name, pool = sys.argv[1], int(sys.argv[2])
    depth, w, h, n, seed, _pool, opts, _ = WORKLOADS[name]
    if po.ref_available(depth):
        secs, _types, stats = po.ref_run_lookahead(depth, w, h, n, seed, opts, pool)
        out = {"seconds": secs, "threads": pool, "kind": "reference", "frames": n, "stats": stats}
    else:
        t = po.Trace(os.path.join(ROOT, "tests", "golden", name + ".trace"))
        r = po.OracleReplay(t)
        t0 = time.time()
        r.run()
        out = {"seconds": time.time() - t0, "threads": 1, "kind": "port", "frames": n}
    print(json.dumps(out))
    sys.stdout.flush()
    os._exit(0)


if __name__ == "__main__":
    main()
