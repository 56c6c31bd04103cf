"""x265 1.9's own lookahead on a named workload, as a child process of bench.py (either arm).

usage: refrun.py <workload> <pool threads> [runs] [gpu|ref] [trace path]
  ->  one JSON line {"seconds": [...], "threads", "kind", "frames", "stats", "totals"}

ref: oracle/_ref/libx265ref<depth>.so -- the UNMODIFIED x265 1.9 Lookahead class driven frame by frame (no frame
     encoders) on the host cores, compiled from /root/reference by oracle/build_ref.py (kind "reference").
     kind "port": only if that library is missing -- the single-threaded oracle replaying the trace.
gpu: oracle/_ref/libx265gpu<depth>.so -- the same x265 objects with the lookahead bound to libx265cu.so (integration/).
The process leaves with os._exit: the reference's static destructors do not survive interpreter teardown reliably.
"""
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from harness.workloads import WORKLOADS  # noqa: E402
from harness import x265host as xh  # noqa: E402


def main():
    name, pool = sys.argv[1], int(sys.argv[2])
    runs = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    gpu = (sys.argv[4] if len(sys.argv) > 4 else "ref") == "gpu"
    trace = sys.argv[5] if len(sys.argv) > 5 else None
    depth, w, h, n, seed, _pool, opts, _ = WORKLOADS[name]
    if xh.available(depth, gpu):
        d = xh.LaDriver(depth, w, h, n, seed, opts, pool, gpu)
        secs, stats = [], None
        for i in range(runs):
            s, _types, stats = d.run(trace=trace if i == 0 else None, level=1)
            secs.append(s)
        out = {"seconds": secs, "threads": pool, "kind": "x265+gpu" if gpu else "reference", "frames": n, "stats": stats}
        if gpu:
            tot = (C.c_longlong * 8)()
            d.L.x265glue_totals(tot)
            out["totals"] = list(tot)
        d.close()
    elif gpu:
        raise SystemExit("refrun: oracle/_ref/libx265gpu%d.so is not built" % depth)
    else:
        from oracle import pyoracle as po
        t = po.Trace(os.path.join(ROOT, "tests", "golden", name + ".trace"))
        r = po.OracleReplay(t)
        t0 = time.time()
        r.run()
        out = {"seconds": [time.time() - t0], "threads": 1, "kind": "port", "frames": n}
    print(json.dumps(out))
    sys.stdout.flush()
    os._exit(0)


if __name__ == "__main__":
    main()
