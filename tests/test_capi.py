"""CPU tests of the C-ABI boundary: the library loads, exports every declared symbol, validates arguments."""
import ctypes
import os

import pytest

from point_unet_b200 import _lib


@pytest.fixture(scope="module")
def L():
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    return _lib.lib()


def test_exports_every_declared_symbol(L):
    names = _lib.declared_symbols()
    assert len(names) >= 7
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/pointunet_b200.h but not exported"


def test_version_and_counters(L):
    assert b"sm_100a" in L.pu_version()
    assert L.pu_launch_count() >= 0
    assert L.pu_last_cuda_error() == 0


def test_knn_workspace_formula(L):
    a = L.pu_knn_workspace_bytes(1, 180000, 180000, 16)
    b = L.pu_knn_workspace_bytes(4, 180000, 180000, 16)
    assert 0 < a < b < 2 << 30
    assert L.pu_knn_workspace_bytes(0, 10, 10, 1) == 0


def test_knn_argument_validation_without_gpu(L):
    # argument errors are reported before any CUDA call, so this runs on a CPU-only box
    assert L.pu_knn_batch(None, None, 1, 10, 10, 16, None, None, 0, None) == -1
    buf = ctypes.create_string_buffer(64)
    p = ctypes.cast(buf, ctypes.c_void_p)
    assert L.pu_knn_batch(p, p, 1, 10, 10, 0, p, None, 0, None) == -1      # K < 1
    assert L.pu_knn_batch(p, p, 1, 10, 10, 33, p, None, 0, None) == -1     # K > PU_KNN_MAX_K
    assert L.pu_knn_batch(p, p, 1, 10, 10, 16, p, None, 0, None) == -2     # no workspace
    assert L.pu_knn_batch(p, p, 0, 10, 10, 16, p, None, 0, None) == 0      # empty batch is a no-op
    assert L.pu_knn_batch(p, p, 1, 10, 0, 16, p, None, 0, None) == 0       # no queries is a no-op


def test_att16_and_inverse_argument_validation_without_gpu(L):
    """The d = 16 attentive-pooling entries and the inverse-list builder reject bad arguments before any CUDA call."""
    from point_unet_b200 import ops
    ops._L()  # declares argument types
    assert L.pu_att16_supported(16, 16, 16) == 1 and L.pu_att16_supported(16, 16, 32) == 1
    assert L.pu_att16_supported(16, 32, 32) == 0 and L.pu_att16_supported(8, 16, 16) == 0 and L.pu_att16_supported(16, 16, 18) == 0
    small, big = L.pu_att16_workspace_bytes(10), L.pu_att16_workspace_bytes(720000)
    assert 0 < small <= big <= 2 * 148 * 256 * 4 + 256            # one [16,16] partial per CTA, at most 2 CTAs per SM
    buf = ctypes.create_string_buffer(4096)
    p = ctypes.cast(buf, ctypes.c_void_p).value
    p = (p + 15) & ~15
    assert L.pu_att16_fwd(None, 16, p, 4, p, 16, None) == -1       # null tensor
    assert L.pu_att16_fwd(p, 8, p, 4, p, 16, None) == -1           # row stride below the channel count
    assert L.pu_att16_fwd(p, 16, p, 0, p, 16, None) == 0           # no points: no-op
    assert L.pu_att16_bwd(p, 16, p, p, 16, 4, p, 16, p, 0, None, 0, None) == -2   # no workspace
    assert L.pu_att16_bwd(p, 16, p, p, 16, 4, p, 18, p, 0, p, 1 << 20, None) == -1  # dx stride not a multiple of 4
    assert L.pu_build_inverse(None, 16, 1, 4, p, p, p, 1 << 20, None) == -1
    assert L.pu_build_inverse(p, 16, 1, 4, p, p, None, 0, None) == -2
    assert L.pu_inverse_workspace_bytes(4, 180000 * 16) > 4 * 4 * 180000 * 16


def test_no_cpu_fallback():
    import numpy as np
    import torch
    from point_unet_b200.helper_tool import DataProcessing
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.PointUnetError):
        DataProcessing.knn_search(np.zeros((1, 8, 3), np.float32), np.zeros((1, 8, 3), np.float32), 2)
