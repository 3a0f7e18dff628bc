"""CPU-side contract tests: the C-ABI library loads and exports every symbol include/rtm3d_decode.h declares, the
ctypes table mirrors the header, argument validation works without a GPU, and the product has no CPU fallback."""
import ctypes
import os
import re

import pytest
import torch

import rtm3d_b200
from rtm3d_b200 import _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "rtm3d_decode.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    decls = {}
    for m in re.finditer(r"\b(?:int|const char\*)\s+(rtm3d_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        args = [a.strip() for a in m.group(2).replace("\n", " ").split(",")]
        decls[m.group(1)] = [] if args == ["void"] else args
    return decls


def test_library_exports_every_declared_symbol():
    decls = _declared()
    assert len(decls) >= 10
    lib = ctypes.CDLL(_native.LIB_PATH)
    for name in decls:
        assert hasattr(lib, name), f"{name} declared in include/rtm3d_decode.h but not exported"


def test_ctypes_table_matches_header():
    decls = _declared()
    assert set(decls) == set(_native.SIGNATURES), set(decls) ^ set(_native.SIGNATURES)
    for name, args in decls.items():
        assert len(args) == len(_native.SIGNATURES[name]), f"{name}: header has {len(args)} parameters"
        for a, t in zip(args, _native.SIGNATURES[name]):
            if "*" in a:
                assert t in (ctypes.c_void_p,) or isinstance(t, type(ctypes.POINTER(ctypes.c_size_t))), (name, a, t)
            elif a.startswith("float"):
                assert t is ctypes.c_float, (name, a)
            elif a.startswith("size_t"):
                assert t is ctypes.c_size_t, (name, a)
            elif a.startswith("unsigned"):
                assert t is ctypes.c_uint, (name, a)
            else:
                assert t is ctypes.c_int, (name, a)


def test_abi_version_and_build_info():
    lib = _native.lib()
    assert lib.rtm3d_abi_version() == _native.ABI_VERSION
    info = lib.rtm3d_build_info().decode()
    assert "sm_100a" in info


def test_argument_validation_needs_no_gpu():
    lib = _native.lib()
    n = ctypes.c_size_t(0)
    assert lib.rtm3d_decode_workspace_bytes(4, 3, 96, 320, 100, ctypes.byref(n)) == 0 and n.value > 0
    assert lib.rtm3d_decode_workspace_bytes(4, 3, 96, 320, 5000, ctypes.byref(n)) == -3       # RTM3D_ERR_TOPK
    assert b"K=5000" in lib.rtm3d_last_error()
    assert lib.rtm3d_decode_workspace_bytes(0, 3, 96, 320, 10, ctypes.byref(n)) == -2          # RTM3D_ERR_SHAPE
    assert lib.rtm3d_decode_workspace_bytes(1, 3, 96, 320, 10, None) == -1                     # RTM3D_ERR_NULL
    # NULL maps are rejected before anything touches the device
    rc = lib.rtm3d_decode_main(None, None, None, 0, 1, 3, 8, 8, 8, 4, 0.4, 4.0, None, None, None, None, None, None, None,
                               None, 0, 0, None)
    assert rc == -1


def test_no_cpu_fallback():
    dec = rtm3d_b200.HeatmapDecoder(0.4, 10, 4.0)
    maps = [torch.zeros(1, c, 8, 8) for c in (3, 16, 2, 2)]
    with pytest.raises(ValueError, match="no CPU path"):
        dec.decode(maps)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "rtm3d_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f"{f} imports the oracle"
