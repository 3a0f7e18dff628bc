"""ctypes binding of librtm3d_decode.so (include/rtm3d_decode.h).  There is NO fallback: if the library is missing or
an entry point fails, the caller gets an exception."""
from __future__ import annotations

import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
# (RTM3D_B200_LIB: developer override, e.g. the `make DEV=1` build with the instrumentation entry points)
LIB_PATH = os.environ.get("RTM3D_B200_LIB") or os.path.join(_HERE, "librtm3d_decode.so")
CSRC = os.path.join(_HERE, "csrc")
ABI_VERSION = 1

F32, BF16 = 0, 1
ERR_SHAPE = -2
FLAG_FORCE_GENERIC = 1
FLAG_NO_SPECULATION = 2
FLAG_NO_GROUP = 4
FLAG_NO_EPILOGUE = 8
FLAG_LEGACY_PLANES = 16
FLAG_NO_SELECT = 32

_c = ctypes
_vp, _i, _f, _sz, _u = _c.c_void_p, _c.c_int, _c.c_float, _c.c_size_t, _c.c_uint

# name -> argtypes, exactly the declarations of include/rtm3d_decode.h
SIGNATURES = {
    "rtm3d_abi_version": [],
    "rtm3d_last_error": [],
    "rtm3d_build_info": [],
    "rtm3d_decode_workspace_bytes": [_i, _i, _i, _i, _i, _c.POINTER(_sz)],
    "rtm3d_workspace_init": [_vp, _sz, _vp],
    "rtm3d_decode_main": [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _f, _f, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _u, _vp],
    "rtm3d_select_main": [_vp, _i, _i, _i, _i, _i, _i, _f, _vp, _vp, _vp, _vp, _sz, _u, _vp],
    "rtm3d_decode_main_host": [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _f, _f, _vp,
                               _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                               _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _u, _vp],
    "rtm3d_decode_keypoints": [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _sz, _u, _vp],
    "rtm3d_decode_keypoints_host": [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _u, _vp],
    "rtm3d_decode_fused": [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _f, _f,
                           _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _u, _vp],
    "rtm3d_decode_fused_gather": [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _f, _f,
                                  _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _u, _vp, _i, _i, _u, _vp],
    "rtm3d_wait_gather": [_vp, _i, _i, _i, _i, _u, _vp],
    "rtm3d_decode_fused_gather_deferred": [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _f, _f,
                                           _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _u, _vp, _vp, _i, _i, _u, _vp],
    "rtm3d_push_gather": [_vp, _i, _i, _i, _i, _i, _u, _vp],
    "rtm3d_signal_gather": [_vp, _i, _i, _i, _i, _i, _u, _vp],
    "rtm3d_epilogue_main": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _f, _vp, _vp, _vp, _vp, _vp],
    "rtm3d_epilogue_keypoints": [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp],
    "rtm3d_decode_fused_host": [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _f, _f, _vp, _vp,
                                _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _u, _vp],
    "rtm3d_post_fused": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _f,
                         _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "rtm3d_select_post": [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _f, _f,
                          _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _u, _vp],
    "rtm3d_group_vertices": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _i, _f, _vp, _vp, _vp, _vp, _vp],
    "rtm3d_fit_box3d": [_vp, _vp, _vp, _vp, _i, _vp, _i, _c.POINTER(_f), _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "rtm3d_encode_main_targets": [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp],
    "rtm3d_gather_l1_loss": [_vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp],
    "rtm3d_gather_l1_loss_grad": [_vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp],
    "rtm3d_focal_loss": [_vp, _vp, _sz, _f, _f, _vp, _vp, _vp],
    "rtm3d_focal_loss_grad": [_vp, _vp, _sz, _f, _f, _vp, _vp, _vp, _vp],
    "rtm3d_pack_wire": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp],
    "rtm3d_sigmoid_f32": [_vp, _vp, _sz, _vp],
    "rtm3d_threshold_table": [_vp, _vp, _i, _c.POINTER(_i), _vp],
    "rtm3d_decode_box3d": [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _f, _f, _vp, _vp, _vp, _vp, _vp, _vp],
}

_lib = None


def build(verbose: bool = False) -> str:
    """Compile the CUDA sources for sm_100a into LIB_PATH (nvcc cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", CSRC], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building librtm3d_decode.so failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stdout)
    return LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `make -C {CSRC}` (or rtm3d_b200.build()). "
                          "rtm3d_b200 has no CPU or PyTorch fallback.")
    L = ctypes.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(L, name)  # AttributeError if the library does not export what the header declares
        fn.argtypes = argtypes
        fn.restype = ctypes.c_char_p if name in ("rtm3d_last_error", "rtm3d_build_info") else ctypes.c_int
    v = L.rtm3d_abi_version()
    if v != ABI_VERSION:
        raise ImportError(f"{LIB_PATH}: ABI version {v}, expected {ABI_VERSION}; rebuild")
    _lib = L
    return L


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().rtm3d_last_error().decode(errors="replace")
        kind = ValueError if rc < 0 else RuntimeError
        raise kind(f"{what} failed (code {rc}): {msg}")
