"""ctypes loader for ``libpointunet_b200.so`` -- the only way the package reaches its CUDA kernels.

There is NO CPU fallback: if the shared library is missing or a call fails, a ``RuntimeError`` is raised.
``build()`` compiles the library in-tree with nvcc for sm_100a (cross-compiles without a GPU).
"""
from __future__ import annotations

import ctypes
import os
import re
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpointunet_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "pointunet_b200.h")

PU_OK = 0
_STATUS = {0: "PU_OK", -1: "PU_ERR_INVALID_ARG", -2: "PU_ERR_WORKSPACE", -3: "PU_ERR_CUDA", -4: "PU_ERR_UNSUPPORTED"}

c_void_p, c_int, c_size_t, c_float = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t, ctypes.c_float


class PointUnetError(RuntimeError):
    pass


def build(verbose: bool = False) -> str:
    """Compile every .cu under csrc/ for sm_100a into the in-tree shared library."""
    proc = subprocess.run(["make", "-s", "-j8", "-C", os.path.join(_HERE, "csrc")], capture_output=True, text=True)
    if proc.returncode != 0:
        raise PointUnetError("nvcc build of libpointunet_b200.so failed:\n" + proc.stdout + proc.stderr)
    if verbose:
        print(proc.stdout + proc.stderr)
    return LIB_PATH


def declared_symbols() -> list[str]:
    """Every function ``include/pointunet_b200.h`` declares (used by the symbol-export test)."""
    text = open(HEADER_PATH).read()
    return sorted(set(re.findall(r"PU_API\s+[\w \*]+?\b(pu_\w+)\s*\(", text)))


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PointUnetError(
                f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback for the PointSegment hot path)")
        L = ctypes.CDLL(LIB_PATH)
        L.pu_version.restype = ctypes.c_char_p
        L.pu_launch_count.restype = ctypes.c_ulonglong
        L.pu_knn_workspace_bytes.restype = c_size_t
        L.pu_knn_workspace_bytes.argtypes = [c_int] * 4
        L.pu_knn_batch.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p]
        L.pu_knn_batch_dist.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                        c_size_t, c_void_p]
        L.pu_knn_self_interp.argtypes = [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                                         c_void_p]
        L.pu_knn_read_stats.argtypes = [c_void_p, ctypes.POINTER(ctypes.c_ulonglong), c_void_p]
        _lib = L
    return _lib


def _declare_ops(L) -> None:
    """argtypes for the LFA-op entry points (all: pointers..., ints..., stream -> int status)."""
    for name, sig in _OP_SIGS.items():
        fn = getattr(L, name, None)
        if fn is None:
            raise PointUnetError(f"libpointunet_b200.so does not export {name}; rebuild it")
        fn.argtypes = sig
        fn.restype = c_int


# filled in by ops.py (keeps the signature next to the Python wrapper that uses it)
_OP_SIGS: dict[str, list] = {}


def check(status: int, what: str) -> None:
    if status != PU_OK:
        extra = ""
        if status == -3:
            extra = f" (cudaError {lib().pu_last_cuda_error()})"
        raise PointUnetError(f"{what} failed: {_STATUS.get(status, status)}{extra}")


def launch_count() -> int:
    return int(lib().pu_launch_count())
