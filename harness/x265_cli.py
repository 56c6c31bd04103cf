"""Run an x265 CLI binary of oracle/_ref (stock `x265_ref<d>` or GPU-lookahead `x265_cu<d>`) on the
synthetic clip and return the md5 of the bitstream (test harness)."""
import hashlib
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import pyoracle as po  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref")

# name -> (depth, width, height, frames, seed, extra CLI options)
CLI_CASES = {
    "cli_720p": (8, 1280, 720, 24, 1234, ["--preset", "medium", "--bframes", "4", "--rc-lookahead", "20"]),
    "cli_360p_b2": (8, 640, 368, 30, 77, ["--preset", "fast", "--bframes", "2", "--rc-lookahead", "12", "--b-adapt", "1"]),
    # weighted bi-prediction: cuTree's bipred weights differ from 32, B slices run the explicit weight analysis on the lowres MVs
    "cli_360p_weightb": (8, 640, 368, 24, 5, ["--preset", "medium", "--bframes", "3", "--rc-lookahead", "15", "--weightb"]),
    # VBV + ABR: cuTreeFinish inside every referenced propagate step, vbvLookahead / frameCostRecalculate reading rowSatds and
    # lowresCosts.  One frame thread: with several, x265's row-level VBV feedback depends on frame-encoder timing.
    "cli_360p_vbv": (8, 640, 368, 24, 5, ["--preset", "medium", "--bframes", "3", "--rc-lookahead", "15", "--bitrate", "1000",
                                          "--vbv-bufsize", "1500", "--vbv-maxrate", "1500", "--frame-threads", "1"]),
    # (VBV together with --weightb is NOT a usable golden: the stock x265 1.9 binary itself writes different bitstreams for that
    # combination depending on thread timing -- three different md5s under `taskset -c 0`, `taskset -c 0,1`, `nice -n 19` --
    # because the B-slice weight analysis of the frame encoders reads lowresMvs[][] fields that vbvLookahead's late estimates
    # are still filling in; each option alone is timing-stable.)
    # the 10-bit build (16-bit pixel kernels) inside the real encoder
    "cli_360p_10bit": (10, 640, 368, 20, 9, ["--preset", "medium", "--bframes", "3", "--rc-lookahead", "12"]),
    # weightb WITHOUT weightp: calcAdaptiveQuantFrame must still produce wp_sum / wp_ssd (slicetype.cpp:138,211)
    "cli_360p_weightb_only": (8, 640, 368, 24, 5, ["--preset", "medium", "--bframes", "3", "--rc-lookahead", "15", "--weightb", "--no-weightp"]),
    # luma AND chroma fade after the scene cut (seed bit 30, oracle/synth.h): weightAnalyse of the frame encoders runs its motion
    # compensation and its weight sweeps on all three planes (on the GPU in the x265_cu build)
    "cli_360p_fade": (8, 640, 368, 30, 0x40000005, ["--preset", "medium", "--bframes", "3", "--rc-lookahead", "15", "--weightb"]),
    "cli_360p_fade10": (10, 640, 368, 24, 0x40000009, ["--preset", "medium", "--bframes", "3", "--rc-lookahead", "12"]),
    # BASELINE.json configs at full size: [0] 1080p x 60 medium / bframes 4 / rc-lookahead 20 (the CLI line BASELINE states),
    # [1] the same clip with --b-adapt 2 --rc-lookahead 40 + cuTree, [2] 4K 8-bit --rc-lookahead 40 --bframes 8 (16 frames),
    # [3] the 10-bit 4K build with --preset slow (12 frames)
    "cli_c0_1080p": (8, 1920, 1080, 60, 1234, ["--preset", "medium", "--bframes", "4", "--rc-lookahead", "20"]),
    "cli_c1_1080p": (8, 1920, 1080, 60, 1234, ["--preset", "medium", "--bframes", "4", "--rc-lookahead", "40", "--b-adapt", "2"]),
    "cli_c2_4k": (8, 3840, 2160, 16, 4321, ["--preset", "medium", "--bframes", "8", "--rc-lookahead", "40", "--b-adapt", "2"]),
    "cli_c3_4k10": (10, 3840, 2160, 12, 4321, ["--preset", "slow"]),
}
# results depend on the pool size: pin it to values every machine can provide
PIN = ["--pools", "4", "--frame-threads", "2"]


def write_clip(path, depth, w, h, n, seed):
    lib = po.oracle(depth)
    import numpy as np
    dt = po.pixel_dtype(depth)
    y = np.zeros((h, w), dt)
    u = np.zeros((h // 2, w // 2), dt)
    v = np.zeros((h // 2, w // 2), dt)
    with open(path, "wb") as f:
        for t in range(n):
            lib.ola_synth_frame(w, h, t, n, seed, y.ctypes.data, w, u.ctypes.data, v.ctypes.data, w // 2)
            f.write(y.tobytes()); f.write(u.tobytes()); f.write(v.tobytes())


def binary(kind, depth):
    return os.path.join(REF, "x265_%s%d" % (kind, depth))


def run_case(kind, name, workdir=None, keep=False):
    depth, w, h, n, seed, opts = CLI_CASES[name]
    exe = binary(kind, depth)
    if not os.path.exists(exe):
        raise FileNotFoundError(exe)
    d = workdir or tempfile.mkdtemp(prefix="x265cli_")
    yuv = os.path.join(d, "%s.yuv" % name)
    if not os.path.exists(yuv):
        write_clip(yuv, depth, w, h, n, seed)
    out = os.path.join(d, "%s_%s.hevc" % (name, kind))
    cmd = [exe, "--input", yuv, "--input-res", "%dx%d" % (w, h), "--fps", "30", "--input-depth", str(depth), "--frames", str(n),
           "--log-level", "error", "--no-progress"] + opts + (PIN[:2] if "--frame-threads" in opts else PIN) + ["-o", out]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=1200)
    if r.returncode != 0:
        raise RuntimeError("%s failed: %s" % (" ".join(cmd), r.stdout[-800:]))
    md5 = hashlib.md5(open(out, "rb").read()).hexdigest()
    size = os.path.getsize(out)
    if not keep:
        os.remove(out)
    return md5, size, d


if __name__ == "__main__":
    import json
    kind = sys.argv[1] if len(sys.argv) > 1 else "ref"
    res = {}
    import time
    only = [a for a in sys.argv[2:] if not a.startswith("--")]
    old = {}
    if kind == "ref" and only:
        old = json.load(open(os.path.join(ROOT, "tests", "golden", "cli_md5.json")))
    res.update(old)
    for nm in CLI_CASES:
        if only and nm not in only:
            continue
        t0 = time.time()
        md5, size, _ = run_case(kind, nm)
        print("%.1f s" % (time.time() - t0), end=" ")
        res[nm] = {"md5": md5, "bytes": size}
        print(nm, md5, size, flush=True)
    if kind == "ref" and "--write-golden" in sys.argv:
        json.dump(res, open(os.path.join(ROOT, "tests", "golden", "cli_md5.json"), "w"), indent=1)
