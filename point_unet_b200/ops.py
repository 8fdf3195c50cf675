"""PyTorch-facing operators of the PointSegment hot path, each a thin autograd shim over the C-ABI.

Tensor layouts are the reference's (channels-last, fp32 features, int32 indices).  Every op runs a
hand-written sm_100a kernel from ``libpointunet_b200.so`` on the current CUDA stream; there is no CPU or
eager-PyTorch fallback (CPU tensors raise).  PyTorch supplies device memory, streams and autograd
bookkeeping only; per-channel vectors of a few hundred floats (batch-norm scale/shift) are the one thing
computed with torch ops.

Reference interfaces mirrored (PointSegment/RandLANet.py): gather_neighbour :377-386,
relative_pos_encoding :337-343, att_pooling :388-401, random_sample :345-360, nearest_interpolation :362-375;
helper_tf_util.conv2d / conv2d_transpose (helper_tf_util.py:115-250).
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib
from ._lib import c_float, c_int, c_size_t, c_void_p
from .helper_tool import workspace

c_ll = ctypes.c_longlong

_lib._OP_SIGS.update({
    "pu_gather_rows_fwd": [c_void_p, c_int, c_int, c_void_p, c_ll, c_int, c_void_p, c_int, c_int, c_void_p],
    "pu_build_inverse": [c_void_p, c_ll, c_int, c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p],
    "pu_segment_sum": [c_void_p, c_int, c_void_p, c_void_p, c_ll, c_void_p, c_int, c_int, c_int, c_void_p],
    "pu_relative_pos_encoding_fwd": [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p],
    "pu_random_sample_fwd": [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_int,
                             c_void_p],
    "pu_random_sample_bwd": [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_ll,
                             c_int, c_void_p, c_int, c_int, c_void_p],
    "pu_linear_fwd": [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int, c_ll, c_int, c_int, c_int, c_void_p,
                      c_void_p, c_void_p],
    "pu_tc_linear_fwd": [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int, c_ll, c_int, c_int, c_int, c_void_p,
                         c_void_p, c_int, c_void_p, c_void_p, c_size_t, c_void_p],
    "pu_tc_att_pooling_fwd": [c_void_p, c_int, c_void_p, c_ll, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p,
                              c_size_t, c_void_p],
    "pu_tc_att_pooling_bwd": [c_void_p, c_int, c_void_p, c_void_p, c_int, c_ll, c_int, c_int, c_void_p, c_int, c_void_p,
                              c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p],
    "pu_tc_att_pooling_bwd_fused": [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_ll, c_int, c_int, c_void_p, c_int,
                                    c_void_p, c_int, c_int, c_void_p, c_void_p],
    "pu_tc_wgrad": [c_void_p, c_int, c_void_p, c_int, c_ll, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_void_p, c_size_t,
                    c_void_p, c_void_p],
    "pu_linear_fwd_ex": [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int, c_ll, c_int, c_int, c_int, c_void_p,
                         c_void_p, c_int, c_void_p],
    "pu_tc_linear_fwd_ex": [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int, c_ll, c_int, c_int, c_int, c_void_p,
                            c_void_p, c_int, c_void_p, c_void_p, c_size_t, c_int, c_void_p],
    "pu_bn_act_fwd_ex": [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_float, c_ll,
                         c_int, c_void_p, c_int, c_void_p, c_int, c_void_p],
    "pu_bn_bwd_reduce_ex": [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p, c_float, c_ll, c_int,
                            c_void_p, c_void_p, c_void_p],
    "pu_bn_bwd_apply_ex": [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p, c_float, c_void_p,
                           c_void_p, c_void_p, c_ll, c_int, c_void_p, c_int, c_void_p],
    "pu_stats_finalize": [c_void_p, c_void_p, c_int, c_int, c_int, c_ll, c_void_p, c_void_p, c_void_p],
    "pu_wgrad": [c_void_p, c_int, c_void_p, c_int, c_ll, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_size_t,
                 c_void_p],
    "pu_bn_act_fwd": [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_float, c_ll, c_int,
                      c_void_p, c_int, c_void_p, c_int, c_void_p],
    "pu_act_bwd": [c_void_p, c_int, c_void_p, c_int, c_float, c_ll, c_int, c_void_p, c_int, c_void_p],
    "pu_bn_bwd_reduce": [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_float, c_ll, c_int, c_void_p,
                         c_void_p, c_void_p],
    "pu_bn_bwd_apply": [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_float, c_void_p, c_void_p,
                        c_void_p, c_ll, c_int, c_void_p, c_int, c_void_p],
    "pu_bn_prepare": [c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                      c_void_p, c_float, c_float, c_void_p],
    "pu_bn_finalize_prepare": [c_void_p, c_void_p, c_int, c_int, c_int, c_ll, c_void_p, c_void_p, c_float, c_void_p, c_void_p,
                               c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_float, c_void_p],
    "pu_bn_bwd_coeffs": [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_ll, c_int, c_void_p, c_void_p,
                         c_void_p, c_void_p, c_void_p, c_void_p],
    "pu_point2prod": [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_size_t,
                      c_void_p],
    "pu_point2label": [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                       c_size_t, c_void_p],
    "pu_volume_argmax": [c_void_p, c_ll, c_int, c_int, c_int, c_void_p, c_void_p],
    "pu_att_pooling_fwd": [c_void_p, c_int, c_void_p, c_ll, c_int, c_int, c_void_p, c_int, c_void_p],
    "pu_att_pooling_bwd": [c_void_p, c_int, c_void_p, c_void_p, c_int, c_ll, c_int, c_int, c_void_p, c_int, c_void_p,
                           c_int, c_void_p],
    "pu_att16_fwd": [c_void_p, c_int, c_void_p, c_ll, c_void_p, c_int, c_void_p],
    "pu_att16_fwd_split": [c_void_p, c_int, c_void_p, c_int, c_void_p, c_ll, c_void_p, c_int, c_void_p],
    "pu_att16_bwd_split": [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int, c_ll, c_void_p, c_int, c_void_p, c_int,
                           c_void_p, c_int, c_void_p, c_size_t, c_void_p],
    "pu_locse_pack_xyz": [c_void_p, c_ll, c_void_p, c_void_p],
    "pu_locse_moments": [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p],
    "pu_locse_bn_prepare": [c_void_p, c_ll, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_float, c_int, c_void_p, c_void_p,
                            c_float, c_float, c_void_p, c_void_p],
    "pu_locse_mlp_fwd": [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_float, c_void_p, c_int, c_void_p,
                         c_int, c_void_p],
    "pu_locse_mlp_bwd": [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_float,
                         c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                         c_void_p],
    "pu_att16_bwd": [c_void_p, c_int, c_void_p, c_void_p, c_int, c_ll, c_void_p, c_int, c_void_p, c_int, c_void_p, c_size_t,
                     c_void_p],
})

import os as _os

# Tensor-core policy for the wide contractions (K, N >= 32): 3 = 3xTF32 on tcgen05 (fp32-class accuracy, default),
# 1 = plain TF32 (stated reduced-precision tolerance), 0 = CUDA-core fp32 tiles only.
TC_MODE = int(_os.environ.get("PU_TC_MODE", "3"))
# d = 16 attentive pooling through the dedicated one-pass kernels (att16.cu); 0 = the generic three-pass CUDA-core path
ATT16 = int(_os.environ.get("PU_ATT16", "1")) != 0
# d = 64 attentive pooling backward with the dgrad through the FC fused into the tcgen05 kernel; 0 = separate accumulate GEMM
ATT_BWD_FUSED = int(_os.environ.get("PU_ATT_BWD_FUSED", "1")) != 0
# position branch of building_block (LocSE -> 10->h conv -> BN -> LeakyReLU) as recompute kernels (csrc/locse_mlp.cu);
# 0 = the separate relative_pos_encoding / linear / batch-norm kernels
LOCSE_FUSED = int(_os.environ.get("PU_LOCSE_FUSED", "1")) != 0
# 16-channel level: keep the two halves of building_block's concat as separate contiguous tensors (att16 reads / writes both);
# 0 = one [.., 16] concat buffer whose 32-byte half rows are written and read by different kernels
ATT16_SPLIT = int(_os.environ.get("PU_ATT16_SPLIT", "1")) != 0
# Storage mode of the pre-normalisation activations y (output of every 1x1 conv that feeds a batch norm; kept from the forward
# for the batch-norm backward -- the largest saved tensors of a training step): "fp32" (default, the parity path) or "bf16"
# (opt-in: y is rounded to bfloat16 when stored, arithmetic and batch statistics stay fp32; stated tolerance rel-L2 <= 2e-2 on
# logits and gradients, arg-max agreement >= 99.5 %, tests/test_bf16_storage_gpu.py).
STORAGE_BF16 = _os.environ.get("PU_STORAGE", "fp32").lower() == "bf16"


def set_storage(mode: str) -> None:
    """``"fp32"`` or ``"bf16"``: see STORAGE_BF16 above.  Global, takes effect for subsequently built graphs / steps."""
    global STORAGE_BF16
    if mode not in ("fp32", "bf16"):
        raise ValueError("storage mode must be 'fp32' or 'bf16'")
    STORAGE_BF16 = mode == "bf16"
_tc_error_flag = {}

LEAKY_SLOPE = 0.2  # helper_tf_util.py:169 (alpha is always 0.2, whatever activation_fn was passed)
BN_EPS = 1e-6      # helper_tf_util.py:167
BN_MOMENTUM = 0.99


def _L():
    L = _lib.lib()
    if not getattr(L, "_pu_extra_declared", False):
        _lib._declare_ops(L)
        L.pu_inverse_workspace_bytes.restype = c_size_t
        L.pu_inverse_workspace_bytes.argtypes = [c_int, c_ll]
        L.pu_wgrad_workspace_bytes.restype = c_size_t
        L.pu_wgrad_workspace_bytes.argtypes = [c_ll, c_int, c_int]
        L.pu_point2prod_workspace_bytes.restype = c_size_t
        L.pu_point2prod_workspace_bytes.argtypes = [c_int, c_int, c_int]
        L.pu_tc_linear_supported.argtypes = [c_ll, c_int, c_int, c_int, c_int, c_int]
        L.pu_tc_att_supported.argtypes = [c_int, c_int, c_int]
        L.pu_tc_att_bwd_fused_supported.argtypes = [c_int, c_int, c_int]
        L.pu_tc_workspace_bytes.restype = c_size_t
        L.pu_tc_workspace_bytes.argtypes = [c_int, c_int]
        L.pu_tc_wgrad_supported.argtypes = [c_ll, c_int, c_int, c_int, c_int, c_int]
        L.pu_tc_wgrad_workspace_bytes.restype = c_size_t
        L.pu_tc_wgrad_workspace_bytes.argtypes = [c_ll, c_int, c_int]
        L.pu_linear_row_tiles.argtypes = [c_ll, c_int, c_int]
        L.pu_linear_rows_per_tile.argtypes = [c_ll, c_int, c_int]
        L.pu_bn_bwd_reduce_blocks.argtypes = [c_ll, c_int]
        L.pu_att16_supported.argtypes = [c_int, c_int, c_int]
        L.pu_att16_workspace_bytes.restype = c_size_t
        L.pu_att16_workspace_bytes.argtypes = [c_ll]
        L.pu_locse_mlp_supported.argtypes = [c_int, c_int]
        L.pu_att16_supported_split.argtypes = [c_int, c_int, c_int, c_int]
        L.pu_locse_mlp_workspace_bytes.restype = c_size_t
        L.pu_locse_mlp_workspace_bytes.argtypes = [c_int]
        L._pu_extra_declared = True
    return L


class KernelTimer:
    """CUDA-event timing of selected C-ABI entry points on the stream they are launched on (bench.py's roofline
    leg).  ``with KernelTimer({"pu_att_pooling_fwd"}) as kt: step(); kt.summary()`` -> {name: (launches, ms)}."""
    active = None

    def __init__(self, names):
        self.names, self.events = set(names), []

    def __enter__(self):
        KernelTimer.active = self
        return self

    def __exit__(self, *exc):
        KernelTimer.active = None

    def summary(self):
        torch.cuda.synchronize()
        out = {}
        for name, tag, e0, e1 in self.events:
            n, ms, tags = out.get(name, (0, 0.0, {}))
            dt = e0.elapsed_time(e1)
            t = tags.get(tag, (0, 0.0))
            tags[tag] = (t[0] + 1, t[1] + dt)
            out[name] = (n + 1, ms + dt, tags)
        return out


def _call(name, *args, tag=None):
    fn = getattr(_L(), name)
    kt = KernelTimer.active
    if name.endswith("_ex"):   # same kernel as the plain entry, with an explicit storage dtype: timed under the plain name
        name = name[:-3]
    if kt is not None and name in kt.names:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        st = fn(*args)
        e1.record()
        kt.events.append((name, tag, e0, e1))
    else:
        st = fn(*args)
    if st != 0:
        _lib.check(st, name)


def _stream(t: torch.Tensor):
    return c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.PointUnetError("point_unet_b200 ops run on CUDA tensors only (no CPU fallback)")


def rows(t: torch.Tensor, keep_dtype: bool = False):
    """View ``t[..., C]`` as a row-strided matrix: returns (tensor, R, C, ld).  Copies only if the layout
    cannot be expressed as rows with one constant stride (e.g. a half of a concat buffer is fine).  ``keep_dtype``: a
    bfloat16 tensor stays bfloat16 (bf16 storage mode; the stride is in elements either way)."""
    if t.dtype != torch.float32 and not (keep_dtype and t.dtype == torch.bfloat16):
        t = t.float()
    C = t.shape[-1]
    ok = t.dim() >= 1 and (C == 1 or t.stride(-1) == 1)
    if ok and t.dim() >= 2:
        ld = t.stride(-2) if t.shape[-2] > 1 else max(C, t.stride(-2))
        expect = ld * t.shape[-2]
        for i in range(t.dim() - 3, -1, -1):
            if t.shape[i] > 1 and t.stride(i) != expect:
                ok = False
                break
            expect *= t.shape[i]
        ok = ok and ld >= C
    elif ok:
        ld = C
    if not ok:
        t = t.contiguous()
        ld = C
    R = t.numel() // C if C > 0 else 0
    return t, R, C, ld


def _idx32(idx: torch.Tensor) -> torch.Tensor:
    if idx.dtype != torch.int32:
        idx = idx.to(torch.int32)
    return idx.contiguous()


# ---------------------------------------------------------------------------------------------
# inverse neighbour lists (scatter-free backward)
class InverseIndex:
    """offsets/perm of an index tensor ``idx [B, M, K]`` pointing into ``n_src`` rows per cloud."""

    def __init__(self, idx: torch.Tensor, n_src: int, out=None, build: bool = True):
        """``out = (offsets [B*n_src+1], perm [B*R])`` int32: preallocated storage; ``build=False`` wraps lists that are
        already there (the pipelined training step builds them one step ahead, train.py)."""
        idx = _idx32(idx)
        B = idx.shape[0]
        R = idx[0].numel()
        L = _L()
        self.n_targets = B * n_src
        if out is not None:
            self.offsets, self.perm = out
            assert self.offsets.dtype == torch.int32 and self.offsets.numel() == B * n_src + 1 and self.offsets.is_contiguous()
            assert self.perm.dtype == torch.int32 and self.perm.numel() == B * R and self.perm.is_contiguous()
        else:
            self.offsets = torch.empty(B * n_src + 1, dtype=torch.int32, device=idx.device)
            self.perm = torch.empty(B * R, dtype=torch.int32, device=idx.device)
        if not build:
            return
        nbytes = L.pu_inverse_workspace_bytes(B, R) + 4 * (B * n_src + 1)
        ws = workspace(nbytes, idx.device, slot=1)
        _call("pu_build_inverse", idx.data_ptr(), R, B, n_src, self.offsets.data_ptr(), self.perm.data_ptr(),
                                      ws.data_ptr(), ws.numel(), _stream(idx))


_inverse_cache: dict = {}


def inverse_of(idx: torch.Tensor, n_src: int) -> InverseIndex:
    """Inverse list of ``idx``, cached per index tensor (the pyramid's indices are reused by every layer)."""
    key = (idx.data_ptr(), tuple(idx.shape), idx._version, n_src, idx.device.index)
    inv = _inverse_cache.get(key)
    if inv is None:
        if len(_inverse_cache) > 64:
            _inverse_cache.clear()
        inv = InverseIndex(idx, n_src)
        _inverse_cache[key] = (inv, idx)  # keep idx alive so data_ptr stays unique
        return inv
    return inv[0]


def register_inverse(idx: torch.Tensor, n_src: int, offsets: torch.Tensor, perm: torch.Tensor):
    """Make ``inverse_of(idx, n_src)`` return lists that were built elsewhere (same content as ``idx``)."""
    key = (idx.data_ptr(), tuple(idx.shape), idx._version, n_src, idx.device.index)
    _inverse_cache[key] = (InverseIndex(idx, n_src, out=(offsets, perm), build=False), idx)


def clear_caches():
    _inverse_cache.clear()
    _pre_grads.clear()


def drop_inverse(ptr: int):
    """Forget cached inverse lists of the index buffer at ``ptr``: kernels that fill an index tensor through its raw
    pointer (``knn_search_cuda(out=...)``) do not bump ``_version``, so a reused buffer would otherwise hit a stale entry."""
    for key in [k for k in _inverse_cache if k[0] == ptr]:
        del _inverse_cache[key]


# ---------------------------------------------------------------------------------------------
def gather_rows(src: torch.Tensor, idx: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """``out[b, r, :] = src[b, idx[b, r], :]`` for ``src [B, n, d]`` and ``idx [B, ...]`` (raw, no autograd)."""
    _need_cuda(src, idx)
    idx = _idx32(idx)
    B, n = src.shape[0], src.shape[1]
    s, _, d, ld_s = rows(src)
    Rpc = idx[0].numel() if B > 0 else 0
    if out is None:
        out = torch.empty(tuple(idx.shape) + (d,), dtype=torch.float32, device=src.device)
    o, Ro, do, ld_o = rows(out)
    assert o.data_ptr() == out.data_ptr() and do == d and Ro == B * Rpc, "gather_rows: bad output view"
    _call("pu_gather_rows_fwd", s.data_ptr(), ld_s, n, idx.data_ptr(), Rpc, B, o.data_ptr(), ld_o, d, _stream(src),
          tag=(B * Rpc, B * n, d))
    return out


def segment_sum(grad_out: torch.Tensor, inv: InverseIndex, d: int, out: torch.Tensor | None = None,
                accumulate: bool = False) -> torch.Tensor:
    g, R, dg, ld_g = rows(grad_out)
    assert dg == d
    if out is None:
        out = torch.empty((inv.n_targets, d), dtype=torch.float32, device=grad_out.device)
    o, Ro, do, ld_o = rows(out)
    assert o.data_ptr() == out.data_ptr() and Ro == inv.n_targets
    _call("pu_segment_sum", g.data_ptr(), ld_g, inv.offsets.data_ptr(), inv.perm.data_ptr(), inv.n_targets,
                                   o.data_ptr(), ld_o, d, int(accumulate), _stream(grad_out), tag=(R, inv.n_targets, d))
    return out


class _GatherFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, src, idx):
        ctx.idx, ctx.shape = idx, src.shape
        return gather_rows(src, idx)

    @staticmethod
    def backward(ctx, grad_out):
        B, n, d = ctx.shape
        inv = inverse_of(ctx.idx, n)
        return segment_sum(grad_out, inv, d).view(B, n, d), None


def gather_neighbour(pc: torch.Tensor, neighbor_idx: torch.Tensor) -> torch.Tensor:
    """RandLANet.py:377-386.  ``pc [B,N,d]``, ``neighbor_idx [B,N,K]`` -> ``[B,N,K,d]``."""
    _need_cuda(pc, neighbor_idx)
    if pc.requires_grad:
        return _GatherFn.apply(pc, neighbor_idx)
    return gather_rows(pc, neighbor_idx)


def nearest_interpolation(feature: torch.Tensor, interp_idx: torch.Tensor) -> torch.Tensor:
    """RandLANet.py:362-375.  ``feature [B,N,1,d]``, ``interp_idx [B,up,1]`` -> ``[B,up,1,d]``."""
    _need_cuda(feature, interp_idx)
    f = feature.squeeze(2)
    out = _GatherFn.apply(f, interp_idx) if f.requires_grad else gather_rows(f, interp_idx)
    return out  # idx [B,up,1] -> [B,up,1,d]


def relative_pos_encoding(xyz: torch.Tensor, neigh_idx: torch.Tensor) -> torch.Tensor:
    """RandLANet.py:337-343.  ``xyz [B,N,3]``, ``neigh_idx [B,N,K]`` -> ``[B,N,K,10]`` (no gradient: xyz is data)."""
    _need_cuda(xyz, neigh_idx)
    xyz = xyz.detach().contiguous().float()
    idx = _idx32(neigh_idx)
    B, N, K = idx.shape
    out = torch.empty((B, N, K, 10), dtype=torch.float32, device=xyz.device)
    _call("pu_relative_pos_encoding_fwd", xyz.data_ptr(), idx.data_ptr(), B, N, K, out.data_ptr(), _stream(xyz))
    return out


class _RandomSampleFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feat, pool_idx):
        idx = _idx32(pool_idx)
        B, n = feat.shape[0], feat.shape[1]
        M, K = idx.shape[1], idx.shape[2]
        f, _, d, ld_f = rows(feat)
        out = torch.empty((B, M, d), dtype=torch.float32, device=feat.device)
        ties = torch.empty((B * M, d), dtype=torch.uint8, device=feat.device)
        _call("pu_random_sample_fwd", f.data_ptr(), ld_f, n, idx.data_ptr(), B, M, K, out.data_ptr(), d,
                                             ties.data_ptr(), d, _stream(feat))
        ctx.save_for_backward(f, out, ties)
        ctx.idx, ctx.dims = pool_idx, (B, n, M, K, d, ld_f)
        return out

    @staticmethod
    def backward(ctx, g_out):
        f, out, ties = ctx.saved_tensors
        B, n, M, K, d, ld_f = ctx.dims
        inv = inverse_of(ctx.idx, n)
        g, _, _, ld_g = rows(g_out)
        g_feat = torch.empty((B, n, d), dtype=torch.float32, device=f.device)
        _call("pu_random_sample_bwd", f.data_ptr(), ld_f, out.data_ptr(), d, ties.data_ptr(), g.data_ptr(), ld_g,
                                             inv.offsets.data_ptr(), inv.perm.data_ptr(), inv.n_targets, K,
                                             g_feat.data_ptr(), d, d, _stream(f))
        return g_feat, None


def random_sample(feature: torch.Tensor, pool_idx: torch.Tensor) -> torch.Tensor:
    """RandLANet.py:345-360.  ``feature [B,N,1,d]``, ``pool_idx [B,N',K]`` -> ``[B,N',1,d]`` (max over K)."""
    _need_cuda(feature, pool_idx)
    return _RandomSampleFn.apply(feature.squeeze(2), pool_idx).unsqueeze(2)


# ---------------------------------------------------------------------------------------------
def tc_error_flag(device) -> torch.Tensor:
    key = device.index if device.index is not None else torch.cuda.current_device()
    if key not in _tc_error_flag:
        _tc_error_flag[key] = torch.zeros(1, dtype=torch.int32, device=device)
    return _tc_error_flag[key]


class PreBN:
    """bf16 storage mode: the pre-normalisation activation of a 1x1 conv as the batch-norm ops receive it.  ``data`` is the
    stored bfloat16 tensor; autograd never sees it (it would cast the fp32 gradient of a bf16 tensor to bf16 and back: two
    extra passes and a rounding of every gradient).  The differentiable link between the conv and the batch norm is
    ``carrier``, a one-element fp32 tensor produced by the conv's autograd node; the real gradient dy travels from the batch
    norm's backward to the conv's backward through ``_pre_grads[key]``."""
    __slots__ = ("data", "carrier", "key", "shape")

    def __init__(self, data, carrier, key):
        self.data, self.carrier, self.key, self.shape = data, carrier, key, data.shape


_pre_grads: dict = {}
_pre_counter = [0]
_zero1_cache: dict = {}


def _zero1(device):
    k = device.index if device.index is not None else torch.cuda.current_device()
    if k not in _zero1_cache:
        _zero1_cache[k] = torch.zeros(1, dtype=torch.float32, device=device)
    return _zero1_cache[k]


def _unwrap_pre(y):
    """(tensor handed to autograd, stored tensor, key or None)"""
    if isinstance(y, PreBN):
        return y.carrier, y.data, y.key
    return y, y, None


class StatPartials:
    """Per-tile (sum, M2) partials of a linear kernel, not yet reduced: ``bn_prepare`` turns them into mean / variance /
    invstd / scale / shift in the same launch (pu_bn_finalize_prepare)."""
    __slots__ = ("ssum", "ssq", "rpt", "rows", "C")

    def __init__(self, ssum, ssq, rpt, rows):
        self.ssum, self.ssq, self.rpt, self.rows, self.C = ssum, ssq, rpt, rows, ssum.shape[1]


def linear_raw(x, w, bias=None, out=None, accumulate=False, want_stats=False, wt=None, tc_mode=None, defer_stats=False):
    """y = x w (+ bias) over rows; optionally the batch-norm mean / biased variance of y. (no autograd)
    ``w`` is [K,N]; ``wt`` (optional) is the same weight stored transposed [N,K] -- the tensor-core path wants the
    K-major form and transposes on the fly (a few hundred KB at most) when only ``w`` is given."""
    xr, M, K, ldx = rows(x)
    if w is not None:
        assert w.dim() == 2 and w.shape[0] == K and w.is_contiguous()
        N = w.shape[1]
    else:
        assert wt.dim() == 2 and wt.shape[1] == K and wt.is_contiguous()
        N = wt.shape[0]
    L = _L()
    mode = TC_MODE if tc_mode is None else tc_mode
    y_bf16 = False
    if out is None:
        # bf16 storage mode: y feeds a batch norm (want_stats) and is produced by a kernel with a bf16 store path
        tc_ok = mode in (1, 3) and M >= 128 and L.pu_tc_linear_supported(M, K, N, ldx, K, N) and xr.data_ptr() % 16 == 0
        y_bf16 = STORAGE_BF16 and want_stats and not accumulate and (N & 3) == 0 and \
            (bool(tc_ok) or L.pu_linear_rows_per_tile(M, K, N) == 2048)
        out = torch.empty(tuple(x.shape[:-1]) + (N,), dtype=torch.bfloat16 if y_bf16 else torch.float32, device=x.device)
    o, Mo, No, ldo = rows(out, keep_dtype=True)
    assert o.data_ptr() == out.data_ptr() and Mo == M and No == N and (y_bf16 or o.dtype == torch.float32)
    use_tc = mode in (1, 3) and M >= 128 and L.pu_tc_linear_supported(M, K, N, ldx, K, ldo) and xr.data_ptr() % 16 == 0
    ssum = ssq = None
    if want_stats:
        rpt = 128 if use_tc else L.pu_linear_rows_per_tile(M, K, N)
        tiles = (M + rpt - 1) // rpt
        ssum = torch.empty((tiles, N), dtype=torch.float32, device=x.device)
        ssq = torch.empty((tiles, N), dtype=torch.float32, device=x.device)
    bptr = bias.data_ptr() if bias is not None else None
    if use_tc:
        if wt is None:
            wt = w.t().contiguous()
        tws = workspace(L.pu_tc_workspace_bytes(K, N), x.device, slot=4)
        _call("pu_tc_linear_fwd_ex" if y_bf16 else "pu_tc_linear_fwd", xr.data_ptr(), ldx, wt.data_ptr(), K, bptr, o.data_ptr(), ldo,
              M, K, N, int(accumulate), ssum.data_ptr() if want_stats else None, ssq.data_ptr() if want_stats else None, mode,
              tc_error_flag(x.device).data_ptr(), tws.data_ptr(), tws.numel(), *((1,) if y_bf16 else ()), _stream(x),
              tag=(M, K, N, int(accumulate)))
    else:
        if w is None:
            w = wt.t().contiguous()
        _call("pu_linear_fwd_ex" if y_bf16 else "pu_linear_fwd", xr.data_ptr(), ldx, w.data_ptr(), N, bptr, o.data_ptr(), ldo, M, K,
              N, int(accumulate), ssum.data_ptr() if want_stats else None, ssq.data_ptr() if want_stats else None,
              *((1,) if y_bf16 else ()), _stream(x), tag=(M, K, N, int(accumulate)))
    if not want_stats:
        return out
    if defer_stats:
        return out, StatPartials(ssum, ssq, rpt, M), None
    mean = torch.empty(N, dtype=torch.float32, device=x.device)
    var = torch.empty(N, dtype=torch.float32, device=x.device)
    _call("pu_stats_finalize", ssum.data_ptr(), ssq.data_ptr(), ssum.shape[0], rpt, N, M, mean.data_ptr(), var.data_ptr(),
          _stream(x))
    return out, mean, var


def wgrad_raw(x, dy, want_db=False, tc_mode=None, out=None, out_db=None, accumulate=False, ws_slot=2):
    """dw = x^T dy (and db = column sums of dy): tensor cores for wide shapes, CUDA cores otherwise (no autograd).
    ``out`` / ``out_db`` (contiguous fp32) receive the result, added to their content when ``accumulate``."""
    xr, M, K, ldx = rows(x)
    gr, Mg, N, ldg = rows(dy)
    assert M == Mg
    L = _L()
    dw = out if out is not None else torch.empty((K, N), dtype=torch.float32, device=x.device)
    assert dw.is_contiguous() and dw.numel() == K * N
    db = None
    if want_db:
        db = out_db if out_db is not None else torch.empty(N, dtype=torch.float32, device=x.device)
    acc = int(bool(accumulate))
    mode = TC_MODE if tc_mode is None else tc_mode
    if mode in (1, 3) and xr.data_ptr() % 16 == 0 and gr.data_ptr() % 16 == 0 and \
            L.pu_tc_wgrad_supported(M, K, N, ldx, ldg, int(want_db)):
        ws = workspace(L.pu_tc_wgrad_workspace_bytes(M, K, N), x.device, slot=ws_slot)
        _call("pu_tc_wgrad", xr.data_ptr(), ldx, gr.data_ptr(), ldg, M, K, N, dw.data_ptr(),
              db.data_ptr() if want_db else None, acc, mode, ws.data_ptr(), ws.numel(), tc_error_flag(x.device).data_ptr(),
              _stream(x), tag=(M, K, N))
        return dw, db
    nbytes = L.pu_wgrad_workspace_bytes(M, K, N)
    ws = workspace(nbytes, x.device, slot=ws_slot)
    _call("pu_wgrad", xr.data_ptr(), ldx, gr.data_ptr(), ldg, M, K, N, dw.data_ptr(),
          db.data_ptr() if want_db else None, acc, ws.data_ptr(), ws.numel(), _stream(x), tag=(M, K, N))
    return dw, db


# Gradient sink: inside a Trainer step every parameter's ``.grad`` is a view of one flat, pre-zeroed buffer.  With the sink
# on, the backward functions make the kernels add weight / bias / gamma / beta gradients straight into those views and
# return None for them, instead of handing autograd ~190 small tensors to ``add_`` into the same views one by one.
GRAD_SINK = False


def _sink(param):
    """The parameter's .grad view if gradients may be written into it directly, else None."""
    if not GRAD_SINK or param is None or not param.is_leaf or not param.requires_grad:
        return None
    g = param.grad
    if g is None or g.dtype != torch.float32 or not g.is_contiguous() or g.shape != param.shape:
        return None
    return g


# Weight gradients off the critical path: nothing in the backward chain consumes dw / db, only the optimiser does.  With a
# side stream set (Trainer does it for the duration of loss.backward()) and the gradient sink on, a weight-gradient kernel
# is queued on that stream right after the kernel that produced dy and runs next to the dgrad chain on the main stream --
# the deep levels launch grids of 20-90 CTAs on 148 SMs, so there is room.  x and dy are kept alive until wgrad_join().
WGRAD_STREAM = None
_wgrad_keep: list = []


def _wgrad(x, dy, want_db=False, out=None, out_db=None, accumulate=False):
    side = WGRAD_STREAM
    if side is None or out is None or (want_db and out_db is None):
        return wgrad_raw(x, dy, want_db=want_db, out=out, out_db=out_db, accumulate=accumulate)
    side.wait_stream(torch.cuda.current_stream(x.device))
    with torch.cuda.stream(side):
        res = wgrad_raw(x, dy, want_db=want_db, out=out, out_db=out_db, accumulate=accumulate, ws_slot=5)
    _wgrad_keep.append((x, dy))
    return res


def wgrad_join(device=None):
    """Make the current stream wait for the weight gradients queued on the side stream and release their operands."""
    if WGRAD_STREAM is not None:
        torch.cuda.current_stream(device).wait_stream(WGRAD_STREAM)
    _wgrad_keep.clear()


class _LinearFn(torch.autograd.Function):
    """y = x w + b with batch statistics of y as non-differentiable side outputs."""

    @staticmethod
    def forward(ctx, x, w, bias, want_stats, zero_bias_grad):
        w = w.contiguous()
        defer = want_stats == "defer"
        res = linear_raw(x, w, bias, want_stats=bool(want_stats), defer_stats=defer)
        ctx.zero_bias_grad = zero_bias_grad
        y, mean, var = res if want_stats else (res, None, None)
        if defer:   # hand the raw partials through autograd as two plain tensors; ops.linear re-wraps them
            _LinearFn.last_partials = (mean.rpt, mean.rows)
            mean, var = mean.ssum, mean.ssq
        ctx.pre_key = None
        _LinearFn.last_pre = None
        if y.dtype == torch.bfloat16:   # bf16 storage: autograd gets a one-element carrier, see PreBN
            _pre_counter[0] += 1
            ctx.pre_key = _pre_counter[0]
            _LinearFn.last_pre = (y, ctx.pre_key)
            y = torch.empty(1, dtype=torch.float32, device=x.device)
        ctx.save_for_backward(x, w)
        ctx.has_bias = bias is not None
        ctx.x_needs = x.requires_grad
        ctx.w_param, ctx.b_param = w, bias
        if want_stats:
            ctx.mark_non_differentiable(mean, var)
            # otherwise autograd zero-fills "gradients" for mean / var before every backward: 88 tiny fill launches a step
            ctx.set_materialize_grads(False)
            return y, mean, var
        return y

    @staticmethod
    def backward(ctx, dy, *_):
        if ctx.pre_key is not None:
            dy = _pre_grads.pop(ctx.pre_key, None)   # the batch norm's backward left the real gradient here
        if dy is None:
            return None, None, None, None, None
        x, w = ctx.saved_tensors
        gw, gb = _sink(ctx.w_param), _sink(ctx.b_param)
        if ctx.has_bias and ctx.zero_bias_grad:
            # a bias that feeds a training-mode batch norm has an identically zero gradient (the BN backward makes every
            # column of dy sum to zero); return the exact value instead of accumulating rounding noise
            dw, _ = _wgrad(x, dy, want_db=False, out=gw, accumulate=gw is not None)
            db = None if gb is not None else torch.zeros(dy.shape[-1], dtype=torch.float32, device=dy.device)
        else:
            both = gw is not None and (gb is not None or not ctx.has_bias)
            dw, db = _wgrad(x, dy, want_db=ctx.has_bias, out=gw if both else None, out_db=gb if both else None,
                            accumulate=both)
            if both:
                db = None
            gw = gw if both else None
        dx = None
        if ctx.x_needs:
            dx = linear_raw(dy, None, wt=w)  # dx = dy w^T: the K-major form of w^T is w itself
            dx = dx.view(x.shape)
        return dx, (None if gw is not None else dw), db, None, None


def linear(x, w, bias=None, want_stats=False, zero_bias_grad=False, defer_stats=False):
    """``y = x w + b`` over the last axis (1x1 conv / dense); with ``want_stats`` also returns the batch mean and
    biased variance of ``y`` per channel (non-differentiable side outputs consumed by :func:`bn_act`).
    ``zero_bias_grad``: the caller asserts that ``y`` goes straight into a training-mode batch norm, whose backward
    makes the bias gradient identically zero.  ``defer_stats``: return ``(y, StatPartials, None)`` instead -- the
    per-tile partials, reduced later inside :func:`bn_prepare` together with the BN coefficients (one launch less)."""
    _need_cuda(x, w)
    if want_stats and defer_stats:
        y, ssum, ssq = _LinearFn.apply(x, w, bias, "defer", zero_bias_grad)
        rpt, nrows = _LinearFn.last_partials
        if _LinearFn.last_pre is not None:
            y = PreBN(_LinearFn.last_pre[0], y, _LinearFn.last_pre[1])
        return y, StatPartials(ssum, ssq, rpt, nrows), None
    res = _LinearFn.apply(x, w, bias, want_stats, zero_bias_grad)
    if want_stats and _LinearFn.last_pre is not None:
        return (PreBN(_LinearFn.last_pre[0], res[0], _LinearFn.last_pre[1]),) + tuple(res[1:])
    return res


def _bn_act_fwd_raw(y, scale, shift, slope, out=None, y2=None, scale2=None, shift2=None, out2=None):
    yr, R, C, ldy = rows(y, keep_dtype=True)
    if out is None:
        out = torch.empty(y.shape, dtype=torch.float32, device=y.device)
    o, Ro, Co, ldo = rows(out)
    assert o.data_ptr() == out.data_ptr() and Ro == R and Co == C
    if y2 is not None:
        y2r, R2, C2, ldy2 = rows(y2, keep_dtype=True)
        assert R2 == R and C2 == C
    ldo2 = 0
    if out2 is not None:
        o2, R2o, C2o, ldo2 = rows(out2)
        assert o2.data_ptr() == out2.data_ptr() and R2o == R and C2o == C
    _call("pu_bn_act_fwd_ex", yr.data_ptr(), ldy, int(yr.dtype == torch.bfloat16), scale.data_ptr(), shift.data_ptr(),
          y2r.data_ptr() if y2 is not None else None, ldy2 if y2 is not None else 0,
          int(y2 is not None and y2r.dtype == torch.bfloat16),
          scale2.data_ptr() if y2 is not None else None, shift2.data_ptr() if y2 is not None else None, float(slope), R, C,
          o.data_ptr(), ldo, out2.data_ptr() if out2 is not None else None, ldo2, _stream(y))
    return out


def _bn_sink(gamma, beta):
    g, b = _sink(gamma), _sink(beta)
    return (g, b) if g is not None and b is not None else None


def _bn_bwd_raw(dz, y, scale, shift, slope, gamma, mean, invstd, training, dz2=None, sink=None):
    """Gradient of out = lrelu(BN(y)) wrt y, gamma, beta given dout = dz (+ dz2, a second upstream gradient summed on
    the fly inside the kernels)."""
    dzr, R, C, ldd = rows(dz)
    yr, Ry, Cy, ldy = rows(y, keep_dtype=True)
    yb = int(yr.dtype == torch.bfloat16)
    assert R == Ry and C == Cy
    d2ptr, ldd2 = None, 0
    if dz2 is not None:
        d2r, R2, C2, ldd2 = rows(dz2)
        assert R2 == R and C2 == C
        d2ptr = d2r.data_ptr()
    L = _L()
    blocks = L.pu_bn_bwd_reduce_blocks(R, C)
    p1 = torch.empty((blocks, C), dtype=torch.float32, device=y.device)
    p2 = torch.empty((blocks, C), dtype=torch.float32, device=y.device)
    st = _stream(y)
    _call("pu_bn_bwd_reduce_ex", dzr.data_ptr(), ldd, d2ptr, ldd2, yr.data_ptr(), ldy, yb, scale.data_ptr(), shift.data_ptr(),
          float(slope), R, C, p1.data_ptr(), p2.data_ptr(), st)
    co = torch.empty((5, C), dtype=torch.float32, device=y.device)  # dgamma, dbeta, ka, kb, kc
    dgamma, dbeta, ka, kb, kc = co[0], co[1], co[2], co[3], co[4]
    if sink is not None:  # (gamma.grad, beta.grad): each batch norm runs once per step, its gradients are plain writes
        dgamma, dbeta = sink
    _call("pu_bn_bwd_coeffs", p1.data_ptr(), p2.data_ptr(), blocks, C, mean.data_ptr(), invstd.data_ptr(),
          gamma.data_ptr(), R, int(bool(training)), dgamma.data_ptr(), dbeta.data_ptr(), ka.data_ptr(), kb.data_ptr(),
          kc.data_ptr(), st)
    dy = torch.empty(y.shape, dtype=torch.float32, device=y.device)
    _call("pu_bn_bwd_apply_ex", dzr.data_ptr(), ldd, d2ptr, ldd2, yr.data_ptr(), ldy, yb, scale.data_ptr(), shift.data_ptr(),
          float(slope), ka.data_ptr(), kb.data_ptr(), kc.data_ptr(), R, C, dy.data_ptr(), C, st)
    return dy, dgamma, dbeta


class _BNActFn(torch.autograd.Function):
    """out = leaky_relu_slope(BN(y) [+ BN2(y2)]) with batch statistics given as (mean, var) of y (and y2)."""

    @staticmethod
    def forward(ctx, y, mean, var, gamma, beta, slope, training, moving, y2, mean2, var2, gamma2, beta2, moving2, pre, pre2):
        # pre / pre2: the stored bf16 tensors and gradient keys when y / y2 are PreBN carriers (bf16 storage mode)
        ctx.pre_key = ctx.pre_key2 = None
        if pre is not None:
            y, ctx.pre_key = pre
        if pre2 is not None:
            y2, ctx.pre_key2 = pre2
        invstd, scale, shift = bn_prepare(mean, var, gamma, beta, moving if training else None)
        if isinstance(mean, StatPartials):
            mean = shift[0]   # the fused finalize wrote the batch mean there
        two = y2 is not None
        if two:
            invstd2, scale2, shift2 = bn_prepare(mean2, var2, gamma2, beta2, moving2 if training else None)
            if isinstance(mean2, StatPartials):
                mean2 = shift2[0]
            res = _bn_act_fwd_raw(y, scale, shift, slope, y2=y2, scale2=scale2, shift2=shift2)
            ctx.save_for_backward(y, scale, shift, gamma, mean, invstd, y2, scale2, shift2, gamma2, mean2, invstd2, res)
        else:
            res = _bn_act_fwd_raw(y, scale, shift, slope)
            ctx.save_for_backward(y, scale, shift, gamma, mean, invstd)
        ctx.two, ctx.slope, ctx.training = two, slope, training
        ctx.bn_params = (gamma, beta, gamma2, beta2)
        return res

    @staticmethod
    def backward(ctx, dout):
        none5 = (None,) * 5
        if not ctx.two:
            y, scale, shift, gamma, mean, invstd = ctx.saved_tensors
            sk = _bn_sink(ctx.bn_params[0], ctx.bn_params[1])
            dy, dg, db = _bn_bwd_raw(dout, y, scale, shift, ctx.slope, gamma, mean, invstd, ctx.training, sink=sk)
            if sk is not None:
                dg = db = None
            if ctx.pre_key is not None:
                _pre_grads[ctx.pre_key] = dy
                dy = _zero1(dout.device)
            return (dy, None, None, dg, db, None, None, None) + none5 + (None, None, None)
        y, scale, shift, gamma, mean, invstd, y2, scale2, shift2, gamma2, mean2, invstd2, res = ctx.saved_tensors
        # through the activation first (sign of the output), then each BN branch without activation
        dr, R, C, ldd = rows(dout)
        rr, _, _, ldr = rows(res)
        dz = torch.empty(res.shape, dtype=torch.float32, device=res.device)
        _call("pu_act_bwd", dr.data_ptr(), ldd, rr.data_ptr(), ldr, float(ctx.slope), R, C, dz.data_ptr(), C,
                                   _stream(res))
        sk = _bn_sink(ctx.bn_params[0], ctx.bn_params[1])
        sk2 = _bn_sink(ctx.bn_params[2], ctx.bn_params[3])
        dy, dg, db = _bn_bwd_raw(dz, y, scale, shift, 1.0, gamma, mean, invstd, ctx.training, sink=sk)
        dy2, dg2, db2 = _bn_bwd_raw(dz, y2, scale2, shift2, 1.0, gamma2, mean2, invstd2, ctx.training, sink=sk2)
        if sk is not None:
            dg = db = None
        if sk2 is not None:
            dg2 = db2 = None
        if ctx.pre_key is not None:
            _pre_grads[ctx.pre_key] = dy
            dy = _zero1(dout.device)
        if ctx.pre_key2 is not None:
            _pre_grads[ctx.pre_key2] = dy2
            dy2 = _zero1(dout.device)
        return dy, None, None, dg, db, None, None, None, dy2, None, None, dg2, db2, None, None, None


def bn_prepare(mean, var, gamma, beta, moving=None):
    """invstd / scale / shift per channel in one launch; ``moving = (moving_mean, moving_var, unbias)`` also applies
    the momentum-0.99 moving-average update in place (the reference runs it with the step, RandLANet.py:90,163)."""
    mm = mv = None
    unbias = 1.0
    if moving is not None:
        mm, mv, unbias = moving
    if isinstance(mean, StatPartials):   # training mode, statistics still as per-tile partials: finalize + prepare fused
        sp, C = mean, mean.C
        buf = torch.empty((6, C), dtype=torch.float32, device=sp.ssum.device)  # invstd, scale, [mean; beta], mean, var
        _call("pu_bn_finalize_prepare", sp.ssum.data_ptr(), sp.ssq.data_ptr(), sp.ssum.shape[0], sp.rpt, C, sp.rows,
              gamma.data_ptr(), beta.data_ptr(), BN_EPS, buf[4].data_ptr(), buf[5].data_ptr(), buf[0].data_ptr(),
              buf[1].data_ptr(), buf[2].data_ptr(), mm.data_ptr() if mm is not None else None,
              mv.data_ptr() if mv is not None else None, BN_MOMENTUM, float(unbias), _stream(sp.ssum))
        return buf[0], buf[1], buf[2:4]
    C = mean.numel()
    buf = torch.empty((4, C), dtype=torch.float32, device=mean.device)  # invstd, scale, [mean; beta]
    _call("pu_bn_prepare", mean.data_ptr(), var.data_ptr(), gamma.data_ptr(), beta.data_ptr(), BN_EPS, C,
          buf[0].data_ptr(), buf[1].data_ptr(), buf[2].data_ptr(), mm.data_ptr() if mm is not None else None,
          mv.data_ptr() if mv is not None else None, BN_MOMENTUM, float(unbias), _stream(mean))
    return buf[0], buf[1], buf[2:4]


def bn_act(y, mean, var, gamma, beta, slope=LEAKY_SLOPE, training=True, moving=None,
           y2=None, mean2=None, var2=None, gamma2=None, beta2=None, moving2=None):
    """``leaky_relu_slope(BN(y) [+ BN2(y2)])`` with the given per-channel statistics (batch stats in training,
    moving stats at inference); ``slope=1`` means no activation (helper_tf_util.py:166-169).  ``moving`` =
    ``(moving_mean, moving_var, unbias)`` is updated in place in training mode."""
    ya, yd, key = _unwrap_pre(y)
    y2a, y2d, key2 = _unwrap_pre(y2) if y2 is not None else (None, None, None)
    _need_cuda(yd)
    return _BNActFn.apply(ya, mean, var, gamma, beta, slope, training, moving, y2a, mean2, var2, gamma2, beta2, moving2,
                          (yd, key) if key is not None else None, (y2d, key2) if key2 is not None else None)


class _LFAConcatFn(torch.autograd.Function):
    """``concat([gather_neighbour(f_pc, idx), lrelu(BN(y))], -1)`` of building_block (RandLANet.py:326-328, 331-333)
    without a concat copy: the gather writes the left half of the buffer, the BN/activation kernel writes its result
    into the right half (and, if ``need_fxyz``, also into a tensor of its own for the following ``mlp2``)."""

    @staticmethod
    def forward(ctx, f_pc, idx, y, mean, var, gamma, beta, training, moving, need_fxyz, pre):
        ctx.pre_key = None
        if pre is not None:   # bf16 storage mode: y is a PreBN carrier, the stored tensor comes separately
            y, ctx.pre_key = pre
        B, N, K, h = y.shape
        invstd, scale, shift = bn_prepare(mean, var, gamma, beta, moving if training else None)
        if isinstance(mean, StatPartials):
            mean = shift[0]
        buf = torch.empty((B, N, K, 2 * h), dtype=torch.float32, device=y.device)
        gather_rows(f_pc, idx, out=buf[..., :h])
        if need_fxyz:
            f_xyz = _bn_act_fwd_raw(y, scale, shift, LEAKY_SLOPE, out2=buf[..., h:])
        else:
            f_xyz = None
            _bn_act_fwd_raw(y, scale, shift, LEAKY_SLOPE, out=buf[..., h:])
        ctx.save_for_backward(y, scale, shift, gamma, mean, invstd)
        ctx.idx, ctx.dims, ctx.training, ctx.need_fxyz = idx, (B, N, K, h, f_pc.shape[1]), training, need_fxyz
        ctx.bn_params = (gamma, beta)
        if need_fxyz:
            return buf, f_xyz
        return buf

    @staticmethod
    def backward(ctx, d_buf, d_fxyz=None):
        y, scale, shift, gamma, mean, invstd = ctx.saved_tensors
        B, N, K, h, n_src = ctx.dims
        if d_buf is None:
            d_buf = torch.zeros((B, N, K, 2 * h), dtype=torch.float32, device=y.device)
        inv = inverse_of(ctx.idx, n_src)
        d_fpc = segment_sum(d_buf[..., :h], inv, h).view(B, n_src, h)
        sk = _bn_sink(*ctx.bn_params)
        dy, dg, db = _bn_bwd_raw(d_buf[..., h:], y, scale, shift, LEAKY_SLOPE, gamma, mean, invstd, ctx.training,
                                 dz2=d_fxyz if ctx.need_fxyz else None, sink=sk)
        if sk is not None:
            dg = db = None
        if ctx.pre_key is not None:
            _pre_grads[ctx.pre_key] = dy
            dy = _zero1(y.device)
        return d_fpc, None, dy, None, None, dg, db, None, None, None, None


def lfa_concat(f_pc, idx, y, mean, var, gamma, beta, training=True, moving=None, need_fxyz=True):
    """Fused ``concat([gather_neighbour(f_pc, idx), leaky_relu(BN(y))])``; returns ``(concat, f_xyz)`` (f_xyz is None
    when ``need_fxyz`` is False)."""
    ya, yd, key = _unwrap_pre(y)
    _need_cuda(f_pc, idx, yd)
    res = _LFAConcatFn.apply(f_pc, idx, ya, mean, var, gamma, beta, training, moving, need_fxyz,
                             (yd, key) if key is not None else None)
    return res if need_fxyz else (res, None)


class _LocSEMlpConcatFn(torch.autograd.Function):
    """``concat([gather_neighbour(f_pc, idx), lrelu(BN(conv2d(relative_pos_encoding(xyz, idx))))], -1)`` -- the first half of
    building_block (RandLANet.py:323-328) -- through the recompute kernels of csrc/locse_mlp.cu: the LocSE rows, the
    pre-normalisation tensor and its gradient are never stored; the batch statistics come from the 10x10 covariance of the
    LocSE rows and the weight gradient from the closed form given in that file."""

    @staticmethod
    def forward(ctx, f_pc, idx, xyz, w, bias, gamma, beta, training, mm, mv, unbias, update_moving, pre):
        xyz = xyz.detach().contiguous().float()
        idx = _idx32(idx)
        w = w.contiguous()
        B, N, K = idx.shape
        h = w.shape[1]
        dev = xyz.device
        L = _L()
        st = _stream(xyz)
        coef = torch.empty(5 * h + 112, dtype=torch.float32, device=dev)
        if pre is not None:      # (padded cloud, moments) prepared with the index pyramid, off the critical path
            xyz, mom = pre
        else:
            xyz, mom = locse_prepare(xyz, idx, moments=training)
        upd = training and update_moving
        _call("pu_locse_bn_prepare", mom.data_ptr() if training else None, B * N * K, w.data_ptr(), h, bias.data_ptr(),
              gamma.data_ptr(), beta.data_ptr(), BN_EPS, int(bool(training)),
              mm.data_ptr() if (upd or not training) else None, mv.data_ptr() if (upd or not training) else None, BN_MOMENTUM,
              float(unbias), coef.data_ptr(), st)
        f_xyz = torch.empty((B, N, K, h), dtype=torch.float32, device=dev)
        ctx.save_for_backward(xyz, w, coef, gamma, bias)
        ctx.idx, ctx.dims = idx, (B, N, K, h, f_pc.shape[1] if f_pc is not None else 0)
        ctx.params = (w, bias, gamma, beta)
        ctx.training = training
        ctx.concat = f_pc is not None
        ctx.set_materialize_grads(False)
        if f_pc is None:
            # no concat buffer: the result is handed out TWICE (two aliases of one tensor) so that the gradients of its two
            # consumers -- att_pooling_1 and mlp2 -- arrive separately and are summed inside the backward kernel
            _call("pu_locse_mlp_fwd", xyz.data_ptr(), idx.data_ptr(), B, N, K, w.data_ptr(), h, coef.data_ptr(), LEAKY_SLOPE,
                  f_xyz.data_ptr(), h, None, 0, st, tag=(B * N * K, h))
            return f_xyz, f_xyz.view(f_xyz.shape)
        buf = torch.empty((B, N, K, 2 * h), dtype=torch.float32, device=dev)
        gather_rows(f_pc, idx, out=buf[..., :h])
        _call("pu_locse_mlp_fwd", xyz.data_ptr(), idx.data_ptr(), B, N, K, w.data_ptr(), h, coef.data_ptr(), LEAKY_SLOPE,
              buf.data_ptr() + 4 * h, 2 * h, f_xyz.data_ptr(), h, st, tag=(B * N * K, h))
        return buf, f_xyz

    @staticmethod
    def backward(ctx, d_buf, d_fxyz):
        xyz, w, coef, gamma, bias = ctx.saved_tensors
        B, N, K, h, n_src = ctx.dims
        dev = xyz.device
        d_fpc = None
        if ctx.concat:
            if d_buf is None:
                d_buf = torch.zeros((B, N, K, 2 * h), dtype=torch.float32, device=dev)
            inv = inverse_of(ctx.idx, n_src)
            d_fpc = segment_sum(d_buf[..., :h], inv, h).view(B, n_src, h)
            dz, R, _, ldz = rows(d_buf[..., h:])
        else:   # (d_buf, d_fxyz) are the gradients of the two aliases of f_xyz
            if d_buf is None:
                d_buf, d_fxyz = d_fxyz, None
            if d_buf is None:
                d_buf = torch.zeros((B, N, K, h), dtype=torch.float32, device=dev)
            dz, R, _, ldz = rows(d_buf)
        d2ptr, ld2 = None, 0
        if d_fxyz is not None:
            d2, R2, _, ld2 = rows(d_fxyz)
            assert R2 == R
            d2ptr = d2.data_ptr()
        wp, bp, gp, bep = ctx.params
        gw, gb, sk = _sink(wp), _sink(bp), _bn_sink(gp, bep)
        dw = gw if gw is not None else torch.empty((10, h), dtype=torch.float32, device=dev)
        if sk is not None:
            dg, dbeta = sk
        else:
            dg = torch.empty(h, dtype=torch.float32, device=dev)
            dbeta = torch.empty(h, dtype=torch.float32, device=dev)
        # training mode: the bias gradient is identically zero (nothing to add to a sink); with moving statistics it is not
        db = None if (gb is not None and ctx.training) else torch.empty(h, dtype=torch.float32, device=dev)
        ws = workspace(_L().pu_locse_mlp_workspace_bytes(h), dev, slot=6)
        _call("pu_locse_mlp_bwd", xyz.data_ptr(), ctx.idx.data_ptr(), B, N, K, w.data_ptr(), h, coef.data_ptr(), gamma.data_ptr(),
              bias.data_ptr(), int(bool(ctx.training)), LEAKY_SLOPE, dz.data_ptr(), ldz, d2ptr, ld2, dw.data_ptr(), int(gw is not None),
              db.data_ptr() if db is not None else None, dg.data_ptr(), dbeta.data_ptr(), ws.data_ptr(), ws.numel(),
              _stream(xyz), tag=(B * N * K, h))
        if gb is not None and db is not None:
            gb.add_(db)
            db = None
        return (d_fpc, None, None, None if gw is not None else dw, db, None if sk is not None else dg,
                None if sk is not None else dbeta, None, None, None, None, None, None)


def locse_mlp_supported(K: int, h: int) -> bool:
    return LOCSE_FUSED and bool(_L().pu_locse_mlp_supported(int(K), int(h)))


def locse_prepare(xyz, idx, moments=True, out=None):
    """What the fused position branch needs from the index pyramid alone (no weights involved, so it can be computed with
    the pyramid, off the critical path): the padded cloud ``xyz4 [B,N,4]`` and, with ``moments``, the sums / centred second
    moments ``mom [65]`` of the LocSE rows of ``(xyz, idx)``.  ``out = (xyz4, mom)``: preallocated storage.  No autograd."""
    _need_cuda(xyz, idx)
    xyz = xyz.detach().contiguous().float()
    idx = _idx32(idx)
    B, N, K = idx.shape
    dev = xyz.device
    st = _stream(xyz)
    if out is not None:
        xyz4, mom = out
        assert xyz4.is_contiguous() and xyz4.numel() == B * N * 4 and mom.numel() >= 65
    else:
        xyz4 = torch.empty((B, N, 4), dtype=torch.float32, device=dev)
        mom = torch.empty(65, dtype=torch.float32, device=dev) if moments else None
    _call("pu_locse_pack_xyz", xyz.data_ptr(), B * N, xyz4.data_ptr(), st)
    if moments:
        ws = workspace(_L().pu_locse_mlp_workspace_bytes(1), dev, slot=7)
        _call("pu_locse_moments", xyz4.data_ptr(), idx.data_ptr(), B, N, K, mom.data_ptr(), ws.data_ptr(), ws.numel(), st,
              tag=(B * N * K,))
    return xyz4, mom


def locse_mlp_concat(xyz, f_pc, idx, w, bias, gamma, beta, training=True, moving_mean=None, moving_var=None, unbias=1.0,
                     update_moving=False, pre=None):
    """Fused position branch of building_block; returns ``(concat [B,N,K,2h], f_xyz [B,N,K,h])``.  ``moving_mean`` /
    ``moving_var`` are the statistics used at inference and, with ``update_moving``, updated in place in training.
    ``f_pc=None``: no concat -- returns ``(f_xyz, f_xyz')``, two aliases of the same tensor, one per consumer.
    ``pre``: the result of :func:`locse_prepare` for ``(xyz, idx)`` when the caller has it already (build_pyramid)."""
    _need_cuda(xyz, idx, w)
    return _LocSEMlpConcatFn.apply(f_pc, idx, xyz, w, bias, gamma, beta, bool(training), moving_mean, moving_var, float(unbias),
                                   bool(update_moving), pre)


# ---------------------------------------------------------------------------------------------
class _AttPoolFn(torch.autograd.Function):
    """f_agg[p,c] = sum_k x[p,k,c] softmax_k(x[p,k,:] w)[c]   (RandLANet.py:394-398, one fused kernel).
    Three kernels by channel width: d = 16 -> att16 (CUDA cores, backward fused into ONE pass), d >= 32 -> tcgen05,
    anything else -> the generic CUDA-core tile."""

    @staticmethod
    def forward(ctx, feature_set, w):
        w = w.contiguous()
        B, N, K, d = feature_set.shape
        x, R, _, ldx = rows(feature_set)
        out = torch.empty((B, N, 1, d), dtype=torch.float32, device=feature_set.device)
        L = _L()
        use16 = ATT16 and x.data_ptr() % 16 == 0 and bool(L.pu_att16_supported(K, d, ldx))
        use_tc = (not use16) and TC_MODE in (1, 3) and B * N * K >= 128 and x.data_ptr() % 16 == 0 and \
            L.pu_tc_att_supported(K, d, ldx)
        wt = w.t().contiguous() if use_tc else None
        if use16:
            _call("pu_att16_fwd", x.data_ptr(), ldx, w.data_ptr(), B * N, out.data_ptr(), d, _stream(x), tag=(B * N, K, d))
        elif use_tc:
            tws = workspace(L.pu_tc_workspace_bytes(d, d), x.device, slot=4)
            _call("pu_tc_att_pooling_fwd", x.data_ptr(), ldx, wt.data_ptr(), B * N, K, d, out.data_ptr(), d, TC_MODE,
                  tc_error_flag(x.device).data_ptr(), tws.data_ptr(), tws.numel(), _stream(x), tag=(B * N, K, d))
        else:
            _call("pu_att_pooling_fwd", x.data_ptr(), ldx, w.data_ptr(), B * N, K, d, out.data_ptr(), d, _stream(x),
                  tag=(B * N, K, d))
        ctx.save_for_backward(x, w, wt if use_tc else w)
        ctx.dims = (B, N, K, d, ldx, use_tc, use16)
        ctx.w_param = w
        return out

    @staticmethod
    def backward(ctx, g_agg):
        x, w, wt = ctx.saved_tensors
        B, N, K, d, ldx, use_tc, use16 = ctx.dims
        g, _, _, ldg = rows(g_agg)
        dx = torch.empty((B, N, K, d), dtype=torch.float32, device=x.device)
        gw = _sink(ctx.w_param)
        if use16:
            dw = gw if gw is not None else torch.empty((d, d), dtype=torch.float32, device=x.device)
            ws = workspace(_L().pu_att16_workspace_bytes(B * N), x.device, slot=2)
            _call("pu_att16_bwd", x.data_ptr(), ldx, w.data_ptr(), g.data_ptr(), ldg, B * N, dx.data_ptr(), d, dw.data_ptr(),
                  int(gw is not None), ws.data_ptr(), ws.numel(), _stream(x), tag=(B * N, K, d))
            return dx, (None if gw is not None else dw)
        d_act = torch.empty((B * N * K, d), dtype=torch.float32, device=x.device)
        if use_tc and ATT_BWD_FUSED and _L().pu_tc_att_bwd_fused_supported(K, d, ldx):
            # d = 64: dx = g s + d_act w^T leaves the kernel complete (second MMA inside the epilogue)
            _call("pu_tc_att_pooling_bwd_fused", x.data_ptr(), ldx, wt.data_ptr(), w.data_ptr(), g.data_ptr(), ldg, B * N, K, d,
                  d_act.data_ptr(), d, dx.data_ptr(), d, TC_MODE, tc_error_flag(x.device).data_ptr(), _stream(x),
                  tag=(B * N, K, d))
            dw, _ = _wgrad(x, d_act, out=gw, accumulate=gw is not None)
            return dx, (None if gw is not None else dw)
        if use_tc:
            tws = workspace(_L().pu_tc_workspace_bytes(d, d), x.device, slot=4)
            _call("pu_tc_att_pooling_bwd", x.data_ptr(), ldx, wt.data_ptr(), g.data_ptr(), ldg, B * N, K, d,
                  d_act.data_ptr(), d, dx.data_ptr(), d, TC_MODE, tc_error_flag(x.device).data_ptr(), tws.data_ptr(),
                  tws.numel(), _stream(x), tag=(B * N, K, d))
        else:
            _call("pu_att_pooling_bwd", x.data_ptr(), ldx, w.data_ptr(), g.data_ptr(), ldg, B * N, K, d,
                  d_act.data_ptr(), d, dx.data_ptr(), d, _stream(x), tag=(B * N, K, d))
        dw, _ = _wgrad(x, d_act, out=gw, accumulate=gw is not None)
        linear_raw(d_act, None, wt=w, out=dx.view(B * N * K, d), accumulate=True)  # dx += d_act w^T
        return dx, (None if gw is not None else dw)


class _AttPoolSplitFn(torch.autograd.Function):
    """:class:`_AttPoolFn` for d = 16 with the two halves of the feature set in separate tensors (att16 kernels)."""

    @staticmethod
    def forward(ctx, left, right, w):
        w = w.contiguous()
        B, N, K, h = left.shape
        xl, R, _, ldl = rows(left)
        xr, R2, _, ldr = rows(right)
        assert R == R2 and right.shape[-1] == h and 2 * h == w.shape[0]
        out = torch.empty((B, N, 1, 2 * h), dtype=torch.float32, device=left.device)
        _call("pu_att16_fwd_split", xl.data_ptr(), ldl, xr.data_ptr(), ldr, w.data_ptr(), B * N, out.data_ptr(), 2 * h, _stream(xl),
              tag=(B * N, K, 2 * h))
        ctx.save_for_backward(xl, xr, w)
        ctx.dims = (B, N, K, h, ldl, ldr)
        ctx.w_param = w
        return out

    @staticmethod
    def backward(ctx, g_agg):
        xl, xr, w = ctx.saved_tensors
        B, N, K, h, ldl, ldr = ctx.dims
        g, _, _, ldg = rows(g_agg)
        dl = torch.empty((B, N, K, h), dtype=torch.float32, device=xl.device)
        dr = torch.empty((B, N, K, h), dtype=torch.float32, device=xl.device)
        gw = _sink(ctx.w_param)
        dw = gw if gw is not None else torch.empty((2 * h, 2 * h), dtype=torch.float32, device=xl.device)
        ws = workspace(_L().pu_att16_workspace_bytes(B * N), xl.device, slot=2)
        _call("pu_att16_bwd_split", xl.data_ptr(), ldl, xr.data_ptr(), ldr, w.data_ptr(), g.data_ptr(), ldg, B * N, dl.data_ptr(), h,
              dr.data_ptr(), h, dw.data_ptr(), int(gw is not None), ws.data_ptr(), ws.numel(), _stream(xl), tag=(B * N, K, 2 * h))
        return dl, dr, (None if gw is not None else dw)


def att_pool_split_supported(K: int, d: int) -> bool:
    return ATT16 and ATT16_SPLIT and bool(_L().pu_att16_supported_split(int(K), int(d), d // 2, d // 2))


def att_pool_split(left: torch.Tensor, right: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    """``att_pool(concat([left, right], -1), w)`` without the concat: ``[B,N,K,8]`` x 2, ``w [16,16]`` -> ``[B,N,1,16]``."""
    _need_cuda(left, right, w)
    if left.data_ptr() % 16 or right.data_ptr() % 16:
        return att_pool(torch.cat([left, right], dim=-1), w)
    return _AttPoolSplitFn.apply(left, right, w)


def att_pool(feature_set: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    """Fused FC + softmax over K + weighted sum: ``[B,N,K,d]`` x ``[d,d]`` -> ``[B,N,1,d]``."""
    _need_cuda(feature_set, w)
    return _AttPoolFn.apply(feature_set, w)


# ---------------------------------------------------------------------------------------------
def point2prod(probs: torch.Tensor, xyz_origin: torch.Tensor, volume_shape, point_idx: torch.Tensor | None = None):
    """testPancreas.py:71-85 / testBraTS.py:83-101 on the device: ``probs [n,C]`` scattered through integer voxel
    coordinates ``xyz_origin [*,3] = (x,y,z)`` into a dense fp32 volume.  ``volume_shape = (Z, X, Y, C)`` is the shape
    the reference allocates; the result has the reference's final layout ``[Z, Y, X, C]`` (after its
    ``np.moveaxis(volume, 1, 2)``).  ``point_idx`` maps point i to its row of ``xyz_origin`` (BraTS)."""
    _need_cuda(probs, xyz_origin)
    Z, X, Y, C = (int(v) for v in volume_shape)
    probs = probs.contiguous().float()
    assert probs.dim() == 2 and probs.shape[1] == C
    xo = xyz_origin.to(torch.int32).contiguous()
    pi = point_idx.to(torch.int32).contiguous() if point_idx is not None else None
    vol = torch.empty((Z, Y, X, C), dtype=torch.float32, device=probs.device)
    ws = workspace(_L().pu_point2prod_workspace_bytes(Z, X, Y), probs.device, slot=3)
    _call("pu_point2prod", probs.data_ptr(), xo.data_ptr(), pi.data_ptr() if pi is not None else None, probs.shape[0], C,
          Z, X, Y, vol.data_ptr(), ws.data_ptr(), ws.numel(), _stream(probs))
    return vol


def point2label(probs: torch.Tensor, xyz_origin: torch.Tensor, volume_shape, point_idx: torch.Tensor | None = None,
                remap=None) -> torch.Tensor:
    """utils/genSegmentationPancreas.py:67-77 / genSegmentationBraTS.py:67-78 fused with ``point2prod``: the uint8 label
    volume ``argmax(prob volume, -1)`` in the layout ``[Z, Y, X]``, written straight from the per-point probabilities
    (the dense probability volume is never materialised).  ``remap=(3, 4)`` is the BraTS relabelling ``seg[seg == 3] = 4``."""
    _need_cuda(probs, xyz_origin)
    Z, X, Y, C = (int(v) for v in volume_shape)
    probs = probs.contiguous().float()
    assert probs.dim() == 2 and probs.shape[1] == C
    xo = xyz_origin.to(torch.int32).contiguous()
    pi = point_idx.to(torch.int32).contiguous() if point_idx is not None else None
    lab = torch.empty((Z, Y, X), dtype=torch.uint8, device=probs.device)
    rf, rt = (int(remap[0]), int(remap[1])) if remap is not None else (-1, 0)
    ws = workspace(_L().pu_point2prod_workspace_bytes(Z, X, Y), probs.device, slot=3)
    _call("pu_point2label", probs.data_ptr(), xo.data_ptr(), pi.data_ptr() if pi is not None else None, probs.shape[0], C,
          Z, X, Y, rf, rt, lab.data_ptr(), ws.data_ptr(), ws.numel(), _stream(probs))
    return lab


def volume_argmax(volume: torch.Tensor, remap=None) -> torch.Tensor:
    """``np.argmax(volume, axis=-1).astype(uint8)`` (+ optional relabelling) of a dense probability volume on the device."""
    _need_cuda(volume)
    v = volume.contiguous().float()
    C = v.shape[-1]
    lab = torch.empty(v.shape[:-1], dtype=torch.uint8, device=v.device)
    rf, rt = (int(remap[0]), int(remap[1])) if remap is not None else (-1, 0)
    _call("pu_volume_argmax", v.data_ptr(), v.numel() // C, C, rf, rt, lab.data_ptr(), _stream(v))
    return lab
