"""``torch.library`` registration of the drop-in boundary (right column).

The five LFA operators of ``PointSegment/RandLANet.py`` and the KNN front-end become dispatcher-visible custom operators

    torch.ops.pointunet.knn_search(support, query, k)                    helper_tool.py:84-94
    torch.ops.pointunet.gather_neighbour(pc, neighbor_idx)               RandLANet.py:377-386
    torch.ops.pointunet.relative_pos_encoding(xyz, neigh_idx)            RandLANet.py:337-343
    torch.ops.pointunet.att_pooling(feature_set, fc_kernel)              RandLANet.py:394-398 (FC + softmax over K + sum)
    torch.ops.pointunet.random_sample(feature, pool_idx)                 RandLANet.py:345-360
    torch.ops.pointunet.nearest_interpolation(feature, interp_idx)       RandLANet.py:362-375

with the reference's argument order and ``[B,N,(K,)d]`` layouts, CUDA-only implementations (the sm_100a kernels behind the
C-ABI; there is no CPU kernel registered, so a CPU tensor fails in the dispatcher), shape functions for meta / fake
tensors, and registered autograd formulas that call the scatter-free backward kernels.  ``Network`` itself keeps calling the
``autograd.Function`` shims in ``ops.py`` directly (same kernels, less dispatcher overhead per call).
"""
from __future__ import annotations

import torch
from torch.library import custom_op

from . import ops
from .helper_tool import knn_search_cuda


@custom_op("pointunet::knn_search", mutates_args=(), device_types="cuda")
def knn_search(support_pts: torch.Tensor, query_pts: torch.Tensor, k: int) -> torch.Tensor:
    return knn_search_cuda(support_pts, query_pts, int(k))


@knn_search.register_fake
def _(support_pts, query_pts, k):
    return support_pts.new_empty((query_pts.shape[0], query_pts.shape[1], k), dtype=torch.int32)


# ---- gather_neighbour -------------------------------------------------------------------------------------------------
@custom_op("pointunet::gather_neighbour", mutates_args=(), device_types="cuda")
def gather_neighbour(pc: torch.Tensor, neighbor_idx: torch.Tensor) -> torch.Tensor:
    return ops.gather_rows(pc, neighbor_idx)


@gather_neighbour.register_fake
def _(pc, neighbor_idx):
    return pc.new_empty(tuple(neighbor_idx.shape) + (pc.shape[-1],), dtype=torch.float32)


@custom_op("pointunet::gather_neighbour_bwd", mutates_args=(), device_types="cuda")
def gather_neighbour_bwd(grad_out: torch.Tensor, neighbor_idx: torch.Tensor, n_src: int) -> torch.Tensor:
    B, d = neighbor_idx.shape[0], grad_out.shape[-1]
    inv = ops.inverse_of(neighbor_idx, n_src)
    return ops.segment_sum(grad_out, inv, d).view(B, n_src, d)


@gather_neighbour_bwd.register_fake
def _(grad_out, neighbor_idx, n_src):
    return grad_out.new_empty((neighbor_idx.shape[0], n_src, grad_out.shape[-1]))


def _gather_setup(ctx, inputs, output):
    pc, idx = inputs
    ctx.save_for_backward(idx)
    ctx.n_src = pc.shape[1]


def _gather_backward(ctx, grad_out):
    (idx,) = ctx.saved_tensors
    return gather_neighbour_bwd(grad_out.contiguous(), idx, ctx.n_src), None


gather_neighbour.register_autograd(_gather_backward, setup_context=_gather_setup)


# ---- relative_pos_encoding (xyz is data: no gradient) -----------------------------------------------------------------
@custom_op("pointunet::relative_pos_encoding", mutates_args=(), device_types="cuda")
def relative_pos_encoding(xyz: torch.Tensor, neigh_idx: torch.Tensor) -> torch.Tensor:
    return ops.relative_pos_encoding(xyz, neigh_idx)


@relative_pos_encoding.register_fake
def _(xyz, neigh_idx):
    return xyz.new_empty(tuple(neigh_idx.shape) + (10,), dtype=torch.float32)


# ---- nearest_interpolation --------------------------------------------------------------------------------------------
@custom_op("pointunet::nearest_interpolation", mutates_args=(), device_types="cuda")
def nearest_interpolation(feature: torch.Tensor, interp_idx: torch.Tensor) -> torch.Tensor:
    return ops.gather_rows(feature.squeeze(2), interp_idx)


@nearest_interpolation.register_fake
def _(feature, interp_idx):
    return feature.new_empty(tuple(interp_idx.shape) + (feature.shape[-1],), dtype=torch.float32)


def _interp_setup(ctx, inputs, output):
    feature, idx = inputs
    ctx.save_for_backward(idx)
    ctx.n_src = feature.shape[1]


def _interp_backward(ctx, grad_out):
    (idx,) = ctx.saved_tensors
    return gather_neighbour_bwd(grad_out.contiguous(), idx, ctx.n_src).unsqueeze(2), None


nearest_interpolation.register_autograd(_interp_backward, setup_context=_interp_setup)


# ---- random_sample (gather + max over K; the gradient splits evenly among exact ties like tf.reduce_max) ---------------
@custom_op("pointunet::random_sample_fwd", mutates_args=(), device_types="cuda")
def _random_sample_fwd(feature: torch.Tensor, pool_idx: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
    feat = feature.squeeze(2)
    idx = pool_idx.to(torch.int32).contiguous()
    B, n = feat.shape[0], feat.shape[1]
    M, K = idx.shape[1], idx.shape[2]
    f, _, d, ld_f = ops.rows(feat)
    out = torch.empty((B, M, 1, d), dtype=torch.float32, device=feat.device)
    ties = torch.empty((B * M, d), dtype=torch.uint8, device=feat.device)
    ops._call("pu_random_sample_fwd", f.data_ptr(), ld_f, n, idx.data_ptr(), B, M, K, out.data_ptr(), d, ties.data_ptr(), d,
              ops._stream(feat))
    return out, ties


@_random_sample_fwd.register_fake
def _(feature, pool_idx):
    B, M, d = pool_idx.shape[0], pool_idx.shape[1], feature.shape[-1]
    return feature.new_empty((B, M, 1, d), dtype=torch.float32), feature.new_empty((B * M, d), dtype=torch.uint8)


@custom_op("pointunet::random_sample_bwd", mutates_args=(), device_types="cuda")
def _random_sample_bwd(grad_out: torch.Tensor, feature: torch.Tensor, out: torch.Tensor, ties: torch.Tensor,
                       pool_idx: torch.Tensor) -> torch.Tensor:
    feat = feature.squeeze(2)
    B, n = feat.shape[0], feat.shape[1]
    K = pool_idx.shape[2]
    f, _, d, ld_f = ops.rows(feat)
    inv = ops.inverse_of(pool_idx, n)
    g, _, _, ld_g = ops.rows(grad_out.contiguous())
    g_feat = torch.empty((B, n, 1, d), dtype=torch.float32, device=f.device)
    ops._call("pu_random_sample_bwd", f.data_ptr(), ld_f, out.data_ptr(), d, ties.data_ptr(), g.data_ptr(), ld_g,
              inv.offsets.data_ptr(), inv.perm.data_ptr(), inv.n_targets, K, g_feat.data_ptr(), d, d, ops._stream(f))
    return g_feat


@_random_sample_bwd.register_fake
def _(grad_out, feature, out, ties, pool_idx):
    return torch.empty_like(feature, dtype=torch.float32)


def _rs_setup(ctx, inputs, output):
    feature, pool_idx = inputs
    out, ties = output
    ctx.save_for_backward(feature, out, ties, pool_idx)


def _rs_backward(ctx, grad_out, grad_ties):
    feature, out, ties, pool_idx = ctx.saved_tensors
    return _random_sample_bwd(grad_out, feature, out, ties, pool_idx), None


_random_sample_fwd.register_autograd(_rs_backward, setup_context=_rs_setup)


def random_sample(feature: torch.Tensor, pool_idx: torch.Tensor) -> torch.Tensor:
    """``[B,N,1,d]``, ``[B,N',K]`` -> ``[B,N',1,d]``; the tie counts of the backward stay internal."""
    return _random_sample_fwd(feature, pool_idx)[0]


# ---- att_pooling: FC (no bias) + softmax over K + weighted sum, one kernel each way ------------------------------------
@custom_op("pointunet::att_pooling", mutates_args=(), device_types="cuda")
def att_pooling(feature_set: torch.Tensor, fc_kernel: torch.Tensor) -> torch.Tensor:
    return ops.att_pool(feature_set.detach(), fc_kernel.detach())


@att_pooling.register_fake
def _(feature_set, fc_kernel):
    B, N, K, d = feature_set.shape
    return feature_set.new_empty((B, N, 1, d), dtype=torch.float32)


@custom_op("pointunet::att_pooling_bwd", mutates_args=(), device_types="cuda")
def _att_pooling_bwd(grad_out: torch.Tensor, feature_set: torch.Tensor, fc_kernel: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
    # the op-level backward rebuilds the autograd node of ops.att_pool (one extra fused forward) and runs its fused backward
    with torch.enable_grad():
        x = feature_set.detach().requires_grad_(True)
        w = fc_kernel.detach().requires_grad_(True)
        out = ops.att_pool(x, w)
    dx, dw = torch.autograd.grad(out, (x, w), grad_out.contiguous())
    return dx, dw


@_att_pooling_bwd.register_fake
def _(grad_out, feature_set, fc_kernel):
    return torch.empty_like(feature_set), torch.empty_like(fc_kernel)


def _att_setup(ctx, inputs, output):
    ctx.save_for_backward(*inputs)


def _att_backward(ctx, grad_out):
    feature_set, fc_kernel = ctx.saved_tensors
    dx, dw = _att_pooling_bwd(grad_out, feature_set, fc_kernel)
    return dx, dw


att_pooling.register_autograd(_att_backward, setup_context=_att_setup)

OP_NAMES = ("knn_search", "gather_neighbour", "relative_pos_encoding", "att_pooling", "random_sample_fwd",
            "nearest_interpolation")
