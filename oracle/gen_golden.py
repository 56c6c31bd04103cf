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

# name -> (depth, w, h, nframes, seed, pool threads, options, dump?)
CASES = {
    # tiny: full arrays dumped, whole-frame estimates only (height < 720 disables coop slices)
    "tiny8": (8, 320, 192, 14, 7, 16, [("preset", "medium"), ("bframes", "3"), ("rc-lookahead", "10")], True),
    "tiny10": (10, 320, 192, 14, 7, 16, [("preset", "medium"), ("bframes", "3"), ("rc-lookahead", "10")], True),
    # odd geometry: lowres width/height not multiples of 8 (rounding, edge CUs), no AQ, no weightp, b-adapt 1
    "odd8": (8, 360, 208, 12, 11, 16, [("preset", "medium"), ("bframes", "2"), ("rc-lookahead", "8"), ("aq-mode", "0"),
                                       ("no-weightp", None), ("b-adapt", "1")], True),
    # BASELINE.json configs[0]: 1080p medium, bframes 4, la 20 (reduced to 720p/30 frames for the CPU suite)
    "c0_720p": (8, 1280, 720, 30, 1234, 16, [("preset", "medium"), ("bframes", "4"), ("rc-lookahead", "20")], False),
    "c0_720p10": (10, 1280, 720, 24, 1234, 16, [("preset", "medium"), ("bframes", "4"), ("rc-lookahead", "20")], False),
    # small pool: batching auto-disables after the first batch (slicetype.cpp:1256,1296) -> sliced searches
    "pool3_720p": (8, 1280, 720, 24, 99, 3, [("preset", "medium"), ("bframes", "3"), ("rc-lookahead", "12")], False),
    # configs[0] full size
    "c0_1080p": (8, 1920, 1080, 60, 1234, 16, [("preset", "medium"), ("bframes", "4"), ("rc-lookahead", "20")], False),
    # configs[1]: 1080p, --b-adapt 2 --rc-lookahead 40 with cuTree  (bench workload)
    "c1_1080p": (8, 1920, 1080, 60, 1234, 16, [("preset", "medium"), ("bframes", "4"), ("rc-lookahead", "40"),
                                                ("b-adapt", "2")], False),
    # configs[2]: 4K, --rc-lookahead 40 --bframes 8
    "c2_4k": (8, 3840, 2160, 48, 4321, 16, [("preset", "medium"), ("bframes", "8"), ("rc-lookahead", "40"),
                                            ("b-adapt", "2")], False),
    # configs[3]: 10-bit 4K, --preset slow
    "c3_4k10": (10, 3840, 2160, 32, 4321, 16, [("preset", "slow")], False),
}


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
