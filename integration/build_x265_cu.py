#!/usr/bin/env python3
"""Build x265 1.9 with its lookahead bound to libx265cu.so (needs /root/reference and a previous oracle/build_ref.py --cli).

  oracle/_ref/libx265gpu<depth>.so   x265's own Lookahead as the HOST of the GPU path + the lookahead-only driver
                                     (harness/x265_la_driver.cpp): `bench.py` e2e arm, trace parity tests
  oracle/_ref/x265_cu<depth>         the full x265 CLI built the same way: bitstream md5 parity tests

Only encoder/slicetype.cpp, encoder/weightPrediction.cpp, common/lowres.cpp and common/picyuv.cpp differ from the stock build: they are compiled from
temporary copies with the call-outs of integration/x265_glue.h inserted (make_gpu_sources.py).  Everything else is the
reference's own objects as built by oracle/build_ref.py.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import build_ref  # noqa: E402

REPLACED = ("encoder_slicetype.o", "common_lowres.o", "common_picyuv.o", "encoder_weightPrediction.o")
OURS = ("slicetype_plain.o", "slicetype_gpu.o", "lowres_gpu.o", "picyuv_gpu.o", "weightpred_gpu.o", "weightpred_open.o", "x265_glue.o", "x265_la_driver_gpu.o")


def main():
    depths = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "8").split(",")]
    if not os.path.isdir(build_ref.REF_ROOT):
        print("build_x265_cu: reference tree not present; nothing to do")
        return 0
    pkg = os.path.join(ROOT, "src", "x265_b200")
    for depth in depths:
        d = os.path.join(build_ref.OUT, "d%d" % depth)
        gen = os.path.join(d, "gen")
        obj = os.path.join(d, "obj")
        F = build_ref.flags(depth, gen) + ["-I" + HERE, "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(pkg, "host")]
        script = os.path.join(HERE, "make_gpu_sources.py")
        build_ref.run([sys.executable, script, build_ref.REF_ROOT, gen])
        glue_h = os.path.join(HERE, "x265_glue.h")
        jobs = []
        for name in ("slicetype_gpu", "lowres_gpu", "picyuv_gpu", "weightpred_gpu"):
            jobs.append((os.path.join(gen, name + ".cpp"), os.path.join(obj, name + ".o"), F, [script, glue_h]))
        F11 = [f for f in F if f != "-std=gnu++98"] + ["-std=gnu++11"]
        jobs.append((os.path.join(HERE, "x265_glue.cpp"), os.path.join(obj, "x265_glue.o"), F11,
                     [glue_h, os.path.join(pkg, "host", "lookahead_cu.h"), os.path.join(ROOT, "include", "x265cu.h")]))
        jobs.append((os.path.join(ROOT, "harness", "x265_la_driver.cpp"), os.path.join(obj, "x265_la_driver_gpu.o"), F,
                     [os.path.join(ROOT, "oracle", "ref_hooks.h"), os.path.join(ROOT, "oracle", "synth.h")]))
        for src, o, fl, deps in jobs:
            if not build_ref.newer(o, src, *deps):
                build_ref.run(["g++"] + fl + ["-c", src, "-o", o])
        ref_objs, cli_objs = [], []
        for fn in sorted(os.listdir(obj)):
            if not fn.endswith(".o") or fn in REPLACED or fn in OURS:
                continue
            (cli_objs if fn.startswith("cli_") else ref_objs).append(os.path.join(obj, fn))
        if not cli_objs:
            raise SystemExit("build_x265_cu: run oracle/build_ref.py --cli first")
        gpu_objs = [os.path.join(obj, n) for n in ("slicetype_gpu.o", "lowres_gpu.o", "picyuv_gpu.o", "weightpred_gpu.o", "x265_glue.o")]
        link = ["-L" + pkg, "-lx265cu_host", "-lx265cu", "-Wl,-rpath," + pkg, "-Wl,-rpath,$ORIGIN/../../src/x265_b200", "-lpthread", "-ldl", "-lm"]
        libs = [os.path.join(pkg, "libx265cu_host.so"), os.path.join(pkg, "libx265cu.so")]
        so = os.path.join(build_ref.OUT, "libx265gpu%d.so" % depth)
        drv = os.path.join(obj, "x265_la_driver_gpu.o")
        if not build_ref.newer(so, *(ref_objs + gpu_objs + [drv] + libs)):
            build_ref.run(["g++", "-shared", "-o", so] + ref_objs + gpu_objs + [drv] + link)
        exe = os.path.join(build_ref.OUT, "x265_cu%d" % depth)
        if not build_ref.newer(exe, *(ref_objs + cli_objs + gpu_objs + libs)):
            build_ref.run(["g++", "-o", exe] + cli_objs + ref_objs + gpu_objs + link)
        print("build_x265_cu: %s, %s ok" % (so, exe))
    return 0


if __name__ == "__main__":
    sys.exit(main())
