"""ctypes bindings for the KNN oracles -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

``knn_reference``  the reference's own ``cpp_knn_batch_omp`` (knn_.cxx:104-135) compiled from
                   /root/reference into ``oracle/_ref/libknn_ref.so`` -- called exactly like
                   ``helper_tool.py:93`` (``knn_batch(..., omp=True)`` then ``astype(int32)``).
``knn_restated``   our C restatement (oracle/knn_oracle.c); ``tie_rule=0`` reproduces nanoflann's
                   first-visited tie behaviour, ``tie_rule=1`` is the canonical (distance, index) rule.
``knn_brute``      exhaustive canonical search, the definition of the tie rule.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ORACLE_SO = os.path.join(_HERE, "liboracle_knn.so")
_REF_SO = os.path.join(_HERE, "_ref", "libknn_ref.so")
_REF_SRC = "/root/reference/PointSegment/utils/nearest_neighbors/knn_.cxx"

_f32p = ctypes.POINTER(ctypes.c_float)
_i64p = ctypes.POINTER(ctypes.c_int64)
_i32p = ctypes.POINTER(ctypes.c_int32)


def build(ref: bool = True) -> None:
    """Compile the restatement, and the reference's own C++ when /root/reference is present."""
    targets = ["all"]
    if ref and os.path.exists(_REF_SRC):
        targets.append("ref")
    subprocess.run(["make", "-s", "-C", _HERE] + targets, check=True)


def have_reference() -> bool:
    return os.path.exists(_REF_SO)


_lib = None
_ref = None


def _oracle():
    global _lib
    if _lib is None:
        if not os.path.exists(_ORACLE_SO):
            build(ref=False)
        _lib = ctypes.CDLL(_ORACLE_SO)
        _lib.pu_oracle_knn_batch.restype = ctypes.c_long
        _lib.pu_oracle_knn_batch.argtypes = [_f32p, ctypes.c_long, ctypes.c_long, _f32p, ctypes.c_long,
                                             ctypes.c_long, _i64p, _f32p, ctypes.c_int, ctypes.c_int]
        _lib.pu_oracle_knn_brute.restype = None
        _lib.pu_oracle_knn_brute.argtypes = [_f32p, ctypes.c_long, ctypes.c_long, _f32p, ctypes.c_long,
                                             ctypes.c_long, _i64p, _f32p]
        _lib.pu_oracle_knn_dists.restype = None
        _lib.pu_oracle_knn_dists.argtypes = [_f32p, ctypes.c_long, ctypes.c_long, _f32p, ctypes.c_long,
                                             ctypes.c_long, _i32p, _f32p]
    return _lib


def _reference():
    global _ref
    if _ref is None:
        if not os.path.exists(_REF_SO):
            raise RuntimeError("oracle/_ref/libknn_ref.so missing: run `make -C oracle ref` where /root/reference exists")
        _ref = ctypes.CDLL(_REF_SO)
        fn = getattr(_ref, "_Z17cpp_knn_batch_ompPKfmmmS0_mmPl")  # cpp_knn_batch_omp, knn_.h:17-19
        fn.restype = None
        fn.argtypes = [_f32p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_size_t, _f32p, ctypes.c_size_t,
                       ctypes.c_size_t, ctypes.POINTER(ctypes.c_long)]
        _ref.knn_batch_omp = fn
    return _ref


def _prep(support, query):
    s = np.ascontiguousarray(support, dtype=np.float32)
    q = np.ascontiguousarray(query, dtype=np.float32)
    assert s.ndim == 3 and q.ndim == 3 and s.shape[2] == 3 and q.shape[2] == 3 and s.shape[0] == q.shape[0]
    return s, q


def knn_reference(support, query, k):
    """helper_tool.py:84-94 on the reference's own compiled C++ (int32 [B,N2,k])."""
    s, q = _prep(support, query)
    B, N1, _ = s.shape
    N2 = q.shape[1]
    out = np.zeros((B, N2, k), dtype=np.int64)  # knn.pyx:93
    _reference().knn_batch_omp(s.ctypes.data_as(_f32p), B, N1, 3, q.ctypes.data_as(_f32p), N2, k,
                               out.ctypes.data_as(ctypes.POINTER(ctypes.c_long)))
    return out.astype(np.int32)  # helper_tool.py:94


def knn_restated(support, query, k, tie_rule=1, return_dist=False, omp=True, return_evals=False):
    s, q = _prep(support, query)
    B, N1, _ = s.shape
    N2 = q.shape[1]
    out = np.zeros((B, N2, k), dtype=np.int64)
    dist = np.zeros((B, N2, k), dtype=np.float32) if return_dist else None
    evals = _oracle().pu_oracle_knn_batch(s.ctypes.data_as(_f32p), B, N1, q.ctypes.data_as(_f32p), N2, k,
                                          out.ctypes.data_as(_i64p),
                                          dist.ctypes.data_as(_f32p) if return_dist else None,
                                          int(tie_rule), int(bool(omp)))
    res = (out.astype(np.int32),)
    if return_dist:
        res += (dist,)
    if return_evals:
        res += (evals,)
    return res if len(res) > 1 else res[0]


def knn_brute(support, query, k, return_dist=False):
    s, q = _prep(support, query)
    assert k <= 64
    B, N1, _ = s.shape
    N2 = q.shape[1]
    out = np.zeros((B, N2, k), dtype=np.int64)
    dist = np.zeros((B, N2, k), dtype=np.float32) if return_dist else None
    _oracle().pu_oracle_knn_brute(s.ctypes.data_as(_f32p), B, N1, q.ctypes.data_as(_f32p), N2, k,
                                  out.ctypes.data_as(_i64p), dist.ctypes.data_as(_f32p) if return_dist else None)
    return (out.astype(np.int32), dist) if return_dist else out.astype(np.int32)


def knn_dists(support, query, idx):
    """fp32 squared distances of the given neighbour lists, reference arithmetic (nanoflann.hpp:343-346)."""
    s, q = _prep(support, query)
    idx = np.ascontiguousarray(idx, dtype=np.int32)
    B, N1, _ = s.shape
    N2, K = idx.shape[1], idx.shape[2]
    out = np.zeros((B, N2, K), dtype=np.float32)
    _oracle().pu_oracle_knn_dists(s.ctypes.data_as(_f32p), B, N1, q.ctypes.data_as(_f32p), N2, K,
                                  idx.ctypes.data_as(_i32p), out.ctypes.data_as(_f32p))
    return out
