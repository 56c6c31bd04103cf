"""The C-ABI library: loads, exports every symbol include/x265cu.h declares, and fails loudly
(no CPU fallback) when no CUDA device is present.  No compute calls here."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    txt = open(os.path.join(ROOT, "include", "x265cu.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(x265cu_[a-z0-9_]+)\s*\(", txt)))


def test_header_and_binding_list_agree(built):
    from src.x265_b200 import abi
    assert header_functions() == sorted(abi.ABI_SYMBOLS)


def test_library_exports_every_declared_symbol(built):
    from src.x265_b200 import abi
    L = abi.lib_cu()
    for name in header_functions():
        assert hasattr(L, name), name
    assert L.x265cu_abi_version() == 2


def test_no_torch_or_cxx_types_in_signatures():
    txt = open(os.path.join(ROOT, "include", "x265cu.h")).read()
    assert 'extern "C"' in txt
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    assert "torch" not in txt and "std::" not in txt and "at::" not in txt and "template" not in txt


def test_open_fails_loudly_without_gpu(built):
    from src.x265_b200 import abi
    L = abi.lib_cu()
    if L.x265cu_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(RuntimeError) as e:
        abi.Lookahead(320, 192)
    assert "no CPU fallback" in str(e.value)
    # raw ABI: error code, NULL ctx, message
    import numpy as np
    lut = np.zeros(4 * 32768 + 1, np.uint16)
    cfg = abi.Config(320, 192, 8, 96, 80, 3, 8, 1, 12, 0, 1, lut.ctypes.data + 2 * 32768 * 2, 0, None, 0)
    ctx = C.c_void_p()
    assert L.x265cu_open(C.byref(cfg), C.byref(ctx)) == -2
    assert not ctx.value
    assert b"no usable CUDA device" in L.x265cu_last_error(None)


def test_bad_config_rejected(built):
    from src.x265_b200 import abi
    L = abi.lib_cu()
    cfg = abi.Config(8, 8, 8, 96, 80, 3, 8, 1, 12, 0, 1, None, 0, None, 0)
    ctx = C.c_void_p()
    assert L.x265cu_open(C.byref(cfg), C.byref(ctx)) == -1


def test_product_never_touches_the_oracle():
    """the product package must not import/link/execute anything under oracle/"""
    pkg = os.path.join(ROOT, "src", "x265_b200")
    for dirpath, _dirs, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle/" not in txt.replace("tests/core_emul.cpp", "") or f == "la_core.h", f
                assert "pyoracle" not in txt, f
                assert "ola_" not in txt, f


def test_every_entry_binds_its_thread_to_the_context_device():
    """Static check of csrc/x265cu.cu: every ABI entry that touches the device calls cudaSetDevice(ctx device) itself or
    delegates to an implementation that does.  (The library links its own CUDA runtime: a host thread that has made no call
    yet -- an x265 pool worker whose first call is x265cu_sync -- would otherwise run on the default device 0 and create a
    context there; on an 8-GPU box with one process per GPU that cost 0.6-1.4 s per process, found in round 2.)"""
    src = open(os.path.join(ROOT, "src", "x265_b200", "csrc", "x265cu.cu")).read()
    no_device_work = {"x265cu_abi_version", "x265cu_device_count", "x265cu_last_error", "x265cu_get_geometry"}
    delegates = {"x265cu_frame_init": "frameInitImpl", "x265cu_frame_init_var": "frameInitImpl", "x265cu_pixelcmp_frames": "x265cu_pixelcmp_planes",
                 "x265cu_host_register": "bindDefaultDevice", "x265cu_host_unregister": "bindDefaultDevice"}
    seen = 0
    for m in re.finditer(r"^(?:int|void|const char\*)\s+(x265cu_[a-z0-9_]+)\s*\(", src, re.M):
        name = m.group(1)
        body = src[m.start():src.find("\n}\n", m.start())]
        if name in no_device_work:
            continue
        seen += 1
        ok = "cudaSetDevice" in body or (name in delegates and delegates[name] in body)
        assert ok, "%s makes device calls without binding the thread to the context's device" % name
    assert seen >= 25
    for impl in ("frameInitImpl", "bindDefaultDevice"):
        i = src.find("int " + impl + "(") if impl == "frameInitImpl" else src.find("void " + impl + "(")
        i = src.find("static int " + impl + "(", src.find("static int " + impl + "(") + 1) if impl == "frameInitImpl" else i
        assert i >= 0 and "cudaSetDevice" in src[i:src.find("\n}\n", i)], impl
