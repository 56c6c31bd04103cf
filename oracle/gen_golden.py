#!/usr/bin/env python3
"""Generate the golden traces/dumps under tests/golden/ by running the UNMODIFIED reference
(oracle/_ref, built by build_ref.py from /root/reference) on the deterministic synthetic clips.

TEST INFRASTRUCTURE.  Run in the build container only (needs /root/reference to have been built);
the outputs are small text/binary fixtures that are committed and travel to the GPU box.

Each trace line carries the reference's own results for one pre-lookahead frame (P) or one
frame-cost estimate (J): sums plus CRC32s of every output array (see oracle/ref_shim.cpp).
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import pyoracle as po  # noqa: E402

GOLD = os.environ.get("X265LA_GOLDEN_DIR") or os.path.join(os.path.dirname(po.HERE), "tests", "golden")

sys.path.insert(0, os.path.dirname(po.HERE))
from harness.workloads import WORKLOADS as CASES  # noqa: E402


def main():
    want = sys.argv[1:] or list(CASES)
    os.makedirs(GOLD, exist_ok=True)
    if len(want) > 1:
        # one process per case: the 8- and 10-bit reference builds export the same C++ symbols
        import subprocess
        for name in want:
            subprocess.run([sys.executable, os.path.abspath(__file__), name], check=True)
        return
    for name in want:
        depth, w, h, n, seed, pool, opts, dump = CASES[name]
        tr = os.path.join(GOLD, name + ".trace")
        dp = os.path.join(GOLD, name + ".dump") if dump else None
        t0 = time.time()
        secs, types, stats = po.ref_run_lookahead(depth, w, h, n, seed, opts, pool, tr, dp)
        print("%-10s %dx%d %d-bit %d frames: lookahead %.2fs (%.1f fps) wall %.1fs  %s  types=%s" %
              (name, w, h, depth, n, secs, n / secs, time.time() - t0, stats, "".join(" IiPbB"[t] if 0 <= t < 6 else "?" for t in types)))
        sys.stdout.flush()
    # the reference's thread pool / static destructors do not survive interpreter teardown
    # reliably; results are on disk, leave without running them
    os._exit(0)


if __name__ == "__main__":
    main()
