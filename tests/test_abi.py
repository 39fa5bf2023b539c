"""The C-ABI library builds, loads, and exports exactly what include/pivlfn.h declares (no GPU needed)."""
import ctypes
import os
import re
import subprocess

import pytest

from pivlfn import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "pivlfn.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    decl = {}
    for m in re.finditer(r"\b(int|long long)\s+(pivlfn_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        args = [a.strip() for a in m.group(3).split(",")]
        if args == ["void"]:
            args = []
        decl[m.group(2)] = (m.group(1), args)
    return decl


def test_header_parses():
    d = _declared()
    assert "pivlfn_corr_nchw" in d and "pivlfn_conv_tc" in d and len(d) >= 17


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    for name in _declared():
        assert hasattr(lib, name), name
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l and "pivlfn_" in l}
    assert set(_declared()) <= exported
    # nothing torch-typed crosses the boundary: the library must not link libtorch / libc10
    ldd = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "torch" not in ldd and "c10" not in ldd


def test_ctypes_prototypes_match_header():
    d = _declared()
    assert set(d) == set(_lib.PROTOTYPES)
    for name, (ret, args) in d.items():
        res, argtypes = _lib.PROTOTYPES[name]
        assert len(args) == len(argtypes), name
        assert res is (ctypes.c_longlong if ret == "long long" else ctypes.c_int), name
        for a, t in zip(args, argtypes):
            if "*" in a:
                assert t in (ctypes.c_void_p, ctypes.POINTER(ctypes.c_float)), (name, a)
            elif a.startswith("float"):
                assert t is ctypes.c_float, (name, a)
            elif a.startswith("long long"):
                assert t is ctypes.c_longlong, (name, a)
            elif a.startswith("double"):
                assert t is ctypes.c_double, (name, a)
            else:
                assert t is ctypes.c_int, (name, a)


def test_no_gpu_calls_without_device():
    lib = _lib.load()
    assert lib.pivlfn_abi_version() >= 1
    assert lib.pivlfn_flow_mean_parts() > 0
    # argument validation happens before any CUDA call
    assert lib.pivlfn_corr_nchw(None, None, None, 1, 1, 1, 1, 1, None) == -1
    assert lib.pivlfn_corr_nchw(1, 1, 1, 1, 4, 8, 8, 3, None) == -1        # stride must be 1 or 2


def test_product_path_never_imports_oracle():
    """The shipped package must not reference oracle/ (it is test infrastructure)."""
    pkg = os.path.join(ROOT, "piv_liteflownet-pytorch_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, fn)).read()
                assert "lfn_oracle" not in txt and "import oracle" not in txt and "from oracle" not in txt, fn
