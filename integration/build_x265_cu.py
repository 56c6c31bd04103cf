#!/usr/bin/env python3
"""Build oracle/_ref/x265_cu<depth>: the UNMODIFIED x265 1.9 CLI + encoder whose lookahead calls
libx265cu.so (INTEGRATION PROOF; needs /root/reference and a previous oracle/build_ref.py --cli).

Only encoder/slicetype.cpp differs from the stock build: it is compiled from a temporary copy with
the three call-outs of integration/x265_glue.h inserted (make_gpu_slicetype.py).  Everything else
is the reference's own objects as built by oracle/build_ref.py.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import build_ref  # noqa: E402


def main():
    depths = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "8").split(",")]
    if not os.path.isdir(build_ref.REF_ROOT):
        print("build_x265_cu: reference tree not present; nothing to do")
        return 0
    pkg = os.path.join(ROOT, "src", "x265_b200")
    for depth in depths:
        d = os.path.join(build_ref.OUT, "d%d" % depth)
        gen = os.path.join(d, "gen")
        F = build_ref.flags(depth, gen) + ["-I" + HERE, "-I" + os.path.join(ROOT, "include")]
        hooked = os.path.join(gen, "slicetype_gpu.cpp")
        build_ref.run([sys.executable, os.path.join(HERE, "make_gpu_slicetype.py"),
                       os.path.join(build_ref.REF_ROOT, "encoder/slicetype.cpp"), hooked])
        o_slice = os.path.join(d, "obj", "slicetype_gpu.o")
        o_glue = os.path.join(d, "obj", "x265_glue.o")
        build_ref.run(["g++"] + F + ["-c", hooked, "-o", o_slice])
        build_ref.run(["g++"] + F + ["-c", os.path.join(HERE, "x265_glue.cpp"), "-o", o_glue])
        objs = []
        for fn in sorted(os.listdir(os.path.join(d, "obj"))):
            if not fn.endswith(".o"):
                continue
            if fn in ("encoder_slicetype.o", "slicetype_plain.o", "slicetype_gpu.o", "x265_glue.o"):
                continue
            objs.append(os.path.join(d, "obj", fn))
        cli = [o for o in objs if os.path.basename(o).startswith("cli_")]
        if not cli:
            raise SystemExit("build_x265_cu: run oracle/build_ref.py --cli first")
        exe = os.path.join(build_ref.OUT, "x265_cu%d" % depth)
        build_ref.run(["g++"] + objs + [o_slice, o_glue, "-L" + pkg, "-lx265cu", "-Wl,-rpath," + pkg, "-Wl,-rpath,$ORIGIN/../../src/x265_b200",
                                       "-lpthread", "-ldl", "-o", exe])
        print("build_x265_cu: %s ok" % exe)
    return 0


if __name__ == "__main__":
    sys.exit(main())
