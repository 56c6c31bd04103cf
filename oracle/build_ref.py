#!/usr/bin/env python3
"""Build the UNMODIFIED x265 1.9 reference (C primitives, no asm) into oracle/_ref/.

TEST INFRASTRUCTURE ONLY.  Nothing under src/ (the product) may link or call what this builds.

The reference CMake project does not configure with the container's CMake 4.x, so
the library sources are compiled where they lie under /root/reference with the flags the CMake
project would have added (source/CMakeLists.txt:162-168,226-237,312-343).  No reference source is
copied into this repository: objects and binaries go to oracle/_ref/ (git-ignored).

Products (per bit depth D in {8,10}):
  oracle/_ref/d<D>/libx265.a        all of common/*.cpp + encoder/*.cpp, unmodified, except that
                                    encoder/slicetype.cpp is compiled from a *hooked* temporary
                                    copy made by make_hooked_slicetype.py (trace/dump call-outs
                                    only; arithmetic untouched; a second, un-hooked object is
                                    kept as slicetype_plain.o to prove the hooks change nothing)
  oracle/_ref/libx265ref<D>.so      shim (ref_shim.cpp): C primitives table; + the lookahead-only driver with its
                                    observation hooks (harness/x265_la_driver.cpp)
  oracle/_ref/x265_ref<D>           the stock CLI (bitstream md5 checks)

Usage: python oracle/build_ref.py [--depths 8,10] [--jobs N] [--cli]
"""
import argparse
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("X265_REF_ROOT", "/root/reference/x265_1.9/source")
OUT = os.path.join(HERE, "_ref")


def flags(depth, gen):
    hbd = 1 if depth > 8 else 0
    return [
        "-O3", "-DNDEBUG", "-std=gnu++98", "-fPIC", "-ffast-math", "-mstackrealign", "-fno-exceptions", "-w",
        "-DX265_ARCH_X86=1", "-DX86_64=1", "-DHAVE_INT_TYPES_H=1", "-D__STDC_LIMIT_MACROS=1",
        "-DHIGH_BIT_DEPTH=%d" % hbd, "-DX265_DEPTH=%d" % depth, "-DEXPORT_C_API=1", "-DX265_NS=x265",
        "-DX265_VERSION=1.9", "-DHAVE_STRTOK_R=1",
        "-I" + gen, "-I" + REF_ROOT, "-I" + REF_ROOT + "/common", "-I" + REF_ROOT + "/encoder",
        "-I" + REF_ROOT + "/input", "-I" + REF_ROOT + "/output", "-I" + HERE,
    ]


def run(cmd):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + "\n")
        raise SystemExit("build_ref: command failed")
    return r.stdout


def newer(dst, *srcs):
    if not os.path.exists(dst):
        return False
    t = os.path.getmtime(dst)
    return all(os.path.getmtime(s) <= t for s in srcs if os.path.exists(s))


def build_depth(depth, jobs, cli):
    d = os.path.join(OUT, "d%d" % depth)
    gen = os.path.join(d, "gen")
    os.makedirs(gen, exist_ok=True)
    os.makedirs(os.path.join(d, "obj"), exist_ok=True)
    cfg = os.path.join(gen, "x265_config.h")
    if not os.path.exists(cfg):
        with open(cfg, "w") as f:
            f.write("#ifndef X265_CONFIG_H\n#define X265_CONFIG_H\n#define X265_BUILD 79\n#endif\n")
    F = flags(depth, gen)

    srcs = []
    for sub in ("common", "encoder"):
        for fn in sorted(os.listdir(os.path.join(REF_ROOT, sub))):
            if fn.endswith(".cpp") and fn != "winxp.cpp":
                srcs.append((sub, fn))

    # hooked copy of slicetype.cpp (call-outs only) -> gen/slicetype_hooked.cpp
    hooked = os.path.join(gen, "slicetype_hooked.cpp")
    hook_script = os.path.join(HERE, "make_hooked_slicetype.py")
    if not newer(hooked, hook_script, os.path.join(REF_ROOT, "encoder/slicetype.cpp")):
        run([sys.executable, hook_script, os.path.join(REF_ROOT, "encoder/slicetype.cpp"), hooked])

    tasks = []
    objs = []
    for sub, fn in srcs:
        src = os.path.join(REF_ROOT, sub, fn)
        obj = os.path.join(d, "obj", "%s_%s.o" % (sub, fn[:-4]))
        if (sub, fn) == ("encoder", "slicetype.cpp"):
            plain = os.path.join(d, "obj", "slicetype_plain.o")
            if not newer(plain, src):
                tasks.append(["g++"] + F + ["-c", src, "-o", plain])
            src = hooked
            if not newer(obj, src, os.path.join(HERE, "ref_hooks.h")):
                tasks.append(["g++"] + F + ["-c", src, "-o", obj])
        elif not newer(obj, src):
            tasks.append(["g++"] + F + ["-c", src, "-o", obj])
        objs.append(obj)
    # weightPrediction.cpp keeps mcLuma / mcChroma / weightCost in an anonymous namespace: a second object is compiled from a
    # temporary copy in which that namespace has a name (and weightAnalyse another one), so that the shim can call the
    # reference's own functions when the oracle's restatement of them is pinned (tests/test_oracle_vs_ref.py)
    wp_src = os.path.join(REF_ROOT, "encoder/weightPrediction.cpp")
    wp_open = os.path.join(gen, "weightpred_open.cpp")
    wp_obj = os.path.join(d, "obj", "weightpred_open.o")
    if not newer(wp_open, wp_src, os.path.abspath(__file__)):
        txt = open(wp_src).read()
        if txt.count("namespace {") != 1 or "namespace X265_NS {\nvoid weightAnalyse" not in txt:
            raise SystemExit("build_ref: weightPrediction.cpp anchors not found")
        txt = txt.replace("namespace {", "namespace wpref {")
        txt = txt.replace("namespace X265_NS {\nvoid weightAnalyse", "using namespace wpref;\nnamespace X265_NS {\nvoid weightAnalyse_wpref_copy")
        open(wp_open, "w").write(txt)
    if not newer(wp_obj, wp_open):
        tasks.append(["g++"] + F + ["-c", wp_open, "-o", wp_obj])
    with ThreadPoolExecutor(jobs) as ex:
        list(ex.map(run, tasks))

    lib = os.path.join(d, "libx265.a")
    if tasks or not os.path.exists(lib):
        if os.path.exists(lib):
            os.remove(lib)
        run(["ar", "rcs", lib] + objs)
    # un-hooked library for the stock CLI
    lib_plain = os.path.join(d, "libx265_plain.a")
    if tasks or not os.path.exists(lib_plain):
        if os.path.exists(lib_plain):
            os.remove(lib_plain)
        plain_objs = [o if not o.endswith("encoder_slicetype.o") else os.path.join(d, "obj", "slicetype_plain.o") for o in objs]
        run(["ar", "rcs", lib_plain] + plain_objs)

    # shim .so: reference primitives + lookahead driver with hooks
    shim_src = os.path.join(HERE, "ref_shim.cpp")
    drv_src = os.path.join(os.path.dirname(HERE), "harness", "x265_la_driver.cpp")   # lookahead-only driver + observation hooks
    shim = os.path.join(OUT, "libx265ref%d.so" % depth)
    deps = [shim_src, drv_src, lib, wp_obj, os.path.join(HERE, "ref_hooks.h"), os.path.join(HERE, "synth.h")]
    if os.path.exists(shim_src) and not newer(shim, *deps):
        run(["g++"] + F + ["-shared", shim_src, drv_src, wp_obj, "-o", shim, "-Wl,--whole-archive", lib, "-Wl,--no-whole-archive",
                            "-lpthread", "-ldl", "-lm"])

    if cli:
        exe = os.path.join(OUT, "x265_ref%d" % depth)
        if not newer(exe, lib_plain):
            cli_objs = []
            cl_tasks = []
            for sub in ("input", "output"):
                for fn in sorted(os.listdir(os.path.join(REF_ROOT, sub))):
                    if fn.endswith(".cpp"):
                        o = os.path.join(d, "obj", "cli_%s_%s.o" % (sub, fn[:-4]))
                        cl_tasks.append(["g++"] + F + ["-c", os.path.join(REF_ROOT, sub, fn), "-o", o])
                        cli_objs.append(o)
            for fn in ("x265.cpp", "x265-extras.cpp"):
                o = os.path.join(d, "obj", "cli_%s.o" % fn[:-4])
                cl_tasks.append(["g++"] + F + ["-c", os.path.join(REF_ROOT, fn), "-o", o])
                cli_objs.append(o)
            with ThreadPoolExecutor(jobs) as ex:
                list(ex.map(run, cl_tasks))
            run(["g++"] + cli_objs + [lib_plain, "-lpthread", "-ldl", "-o", exe])
    return shim


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--depths", default="8,10")
    ap.add_argument("--jobs", type=int, default=os.cpu_count() or 4)
    ap.add_argument("--cli", action="store_true")
    a = ap.parse_args()
    if not os.path.isdir(REF_ROOT):
        print("build_ref: reference tree not present (%s); nothing to do" % REF_ROOT)
        return 0
    os.makedirs(OUT, exist_ok=True)
    for dep in [int(x) for x in a.depths.split(",")]:
        build_depth(dep, a.jobs, a.cli)
        print("build_ref: depth %d ok" % dep)
    return 0


if __name__ == "__main__":
    sys.exit(main())
