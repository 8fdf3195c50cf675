"""oracle/randla_ref.py -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

PyTorch-CPU restatement (fp32 or fp64, dtype follows the inputs) of the reference's TensorFlow graph for the
PointSegment hot path, op for op, materialising every intermediate exactly where the reference does.
Autograd on this restatement provides the reference gradients.

PARITY UNPINNED at the TensorFlow boundary: TensorFlow 1.11 (environment.yml:158-160) is not installable here
and the reference ships no tests or golden vectors for these ops, so this file restates the published
semantics of the TF ops named at each call site; the CUDA path is held to 1e-3 relative against the fp64
run of this restatement.

What follows what (paths relative to /root/reference/PointSegment):
  gather_neighbour        RandLANet.py:377-386   (tf.batch_gather on [B, N*K])
  relative_pos_encoding   RandLANet.py:337-343
  att_pooling             RandLANet.py:388-401
  random_sample           RandLANet.py:345-360
  nearest_interpolation   RandLANet.py:362-375
  building_block          RandLANet.py:323-335
  dilated_res_block       RandLANet.py:314-321
  inference               RandLANet.py:110-152
  get_loss                RandLANet.py:267-274 (+ masking :62-84 with empty ignored_label_inds)
  conv2d                  helper_tf_util.py:115-170   (kernel [1,1,Cin,Cout] stored here as [Cin,Cout])
  conv2d_transpose        helper_tf_util.py:173-250   (kernel [1,1,Cout,Cin] stored as [Cout,Cin]; 1x1, stride 1)
  batch norm              tf.layers.batch_normalization(momentum 0.99, eps 1e-6): batch mean / BIASED variance
                          in training, moving statistics at inference
  dropout                 helper_tf_util.py:553-574   (keep 0.5, inverted scaling; the mask is INJECTED)
  tf_map (index pyramid)  runPancreas.py:124-145
  point2prod              testPancreas.py:71-85, testBraTS.py:83-101 (+ the p_idx expansion testBraTS.py:226-231)
"""
from __future__ import annotations

import numpy as np
import torch

BN_EPS = 1e-6
LEAKY = 0.2


def batch_gather(x, idx):
    """tf.batch_gather(x [B,N,d], idx [B,M]) -> [B,M,d]"""
    B, M = idx.shape
    return torch.gather(x, 1, idx.long().unsqueeze(-1).expand(B, M, x.shape[-1]))


def gather_neighbour(pc, neighbor_idx):
    B, N, K = neighbor_idx.shape[0], neighbor_idx.shape[1], neighbor_idx.shape[2]
    d = pc.shape[2]
    index_input = neighbor_idx.reshape(B, -1)
    features = batch_gather(pc, index_input)
    return features.reshape(B, N, K, d)


def relative_pos_encoding(xyz, neigh_idx):
    neighbor_xyz = gather_neighbour(xyz, neigh_idx)
    xyz_tile = xyz.unsqueeze(2).repeat(1, 1, neigh_idx.shape[-1], 1)
    relative_xyz = xyz_tile - neighbor_xyz
    relative_dis = torch.sqrt(torch.sum(torch.square(relative_xyz), dim=-1, keepdim=True))
    return torch.cat([relative_dis, relative_xyz, xyz_tile, neighbor_xyz], dim=-1)


def random_sample(feature, pool_idx):
    feature = feature.squeeze(2)
    num_neigh = pool_idx.shape[-1]
    d = feature.shape[-1]
    B = pool_idx.shape[0]
    pool_features = batch_gather(feature, pool_idx.reshape(B, -1)).reshape(B, -1, num_neigh, d)
    # tf.reduce_max: forward max; its gradient splits evenly among exact ties -- torch.amax does the same
    return torch.amax(pool_features, dim=2, keepdim=True)


def nearest_interpolation(feature, interp_idx):
    feature = feature.squeeze(2)
    B, up = interp_idx.shape[0], interp_idx.shape[1]
    return batch_gather(feature, interp_idx.reshape(B, up)).unsqueeze(2)


def batch_norm(x, p, name, is_training, momentum_updates=None):
    """tf.layers.batch_normalization over the last axis."""
    gamma, beta = p[name + "/gamma"], p[name + "/beta"]
    if is_training:
        red = tuple(range(x.dim() - 1))
        mean = x.mean(dim=red)
        var = ((x - mean) ** 2).mean(dim=red)  # biased
        if momentum_updates is not None:
            momentum_updates[name] = (mean.detach(), var.detach(), x.numel() // x.shape[-1])
    else:
        mean, var = p[name + "/moving_mean"], p[name + "/moving_variance"]
    return (x - mean) * torch.rsqrt(var + BN_EPS) * gamma + beta


def leaky_relu(x):
    return torch.where(x > 0, x, x * LEAKY)


def conv2d(x, p, scope, bn, is_training, activation=True, upd=None):
    y = x @ p[scope + "/weights"] + p[scope + "/biases"]
    if bn:
        y = batch_norm(y, p, scope + "/bn", is_training, upd)
    if activation:
        y = leaky_relu(y)
    return y


def conv2d_transpose(x, p, scope, is_training, upd=None):
    y = x @ p[scope + "/weights"].t() + p[scope + "/biases"]  # kernel stored [Cout, Cin]
    y = batch_norm(y, p, scope + "/bn", is_training, upd)
    return leaky_relu(y)


def att_pooling(feature_set, p, name, is_training, upd=None, keep=None):
    B, N, K, d = feature_set.shape
    f_reshaped = feature_set.reshape(-1, K, d)
    att_activation = f_reshaped @ p[name + "fc/kernel"]           # tf.layers.dense, no bias
    att_scores = torch.softmax(att_activation, dim=1)             # over the neighbour axis, per channel
    f_agg = (f_reshaped * att_scores).sum(dim=1).reshape(B, N, 1, d)
    if keep is not None:
        keep[name + "f_agg"] = f_agg
    return conv2d(f_agg, p, name + "mlp", True, is_training, True, upd)


def building_block(xyz, feature, neigh_idx, d_out, p, name, is_training, upd=None, keep=None):
    f_xyz = relative_pos_encoding(xyz, neigh_idx)
    f_xyz = conv2d(f_xyz, p, name + "mlp1", True, is_training, True, upd)
    f_neighbours = gather_neighbour(feature.squeeze(2), neigh_idx)
    f_concat = torch.cat([f_neighbours, f_xyz], dim=-1)
    f_pc_agg = att_pooling(f_concat, p, name + "att_pooling_1", is_training, upd, keep)
    f_xyz = conv2d(f_xyz, p, name + "mlp2", True, is_training, True, upd)
    f_neighbours = gather_neighbour(f_pc_agg.squeeze(2), neigh_idx)
    f_concat = torch.cat([f_neighbours, f_xyz], dim=-1)
    return att_pooling(f_concat, p, name + "att_pooling_2", is_training, upd, keep)


def dilated_res_block(feature, xyz, neigh_idx, d_out, p, name, is_training, upd=None, keep=None):
    f_pc = conv2d(feature, p, name + "mlp1", True, is_training, True, upd)
    f_pc = building_block(xyz, f_pc, neigh_idx, d_out, p, name + "LFA", is_training, upd, keep)
    f_pc = conv2d(f_pc, p, name + "mlp2", True, is_training, False, upd)
    shortcut = conv2d(feature, p, name + "shortcut", True, is_training, False, upd)
    return leaky_relu(f_pc + shortcut)


def inference(p, inputs, cfg, is_training, dropout_mask=None, upd=None, keep=None, checkpoint=False):
    """RandLANet.py:110-152.  inputs: dict(xyz=[..5], neigh_idx, sub_idx, interp_idx, features [B,N,F]).
    ``checkpoint``: recompute each encoder block in the backward instead of keeping its [B,N,K,d] intermediates (same
    arithmetic, same results; lets the full-size 4 x 180k fp64 golden run fit in host memory)."""
    feature = inputs["features"] @ p["fc0/kernel"] + p["fc0/bias"]
    feature = leaky_relu(batch_norm(feature, p, "fc0/bn", is_training, upd))
    feature = feature.unsqueeze(2)
    f_encoder_list = []
    for i in range(cfg.num_layers):
        if checkpoint:
            from torch.utils.checkpoint import checkpoint as _ckpt
            f_encoder_i = _ckpt(lambda f, i=i: dilated_res_block(f, inputs["xyz"][i], inputs["neigh_idx"][i], cfg.d_out[i], p,
                                                                 "Encoder_layer_" + str(i), is_training, upd, keep),
                                feature, use_reentrant=False)
        else:
            f_encoder_i = dilated_res_block(feature, inputs["xyz"][i], inputs["neigh_idx"][i], cfg.d_out[i], p,
                                            "Encoder_layer_" + str(i), is_training, upd, keep)
        f_sampled_i = random_sample(f_encoder_i, inputs["sub_idx"][i])
        feature = f_sampled_i
        if i == 0:
            f_encoder_list.append(f_encoder_i)
        f_encoder_list.append(f_sampled_i)
        if keep is not None:
            keep["enc_%d" % i] = f_encoder_i
    feature = conv2d(f_encoder_list[-1], p, "decoder_0", True, is_training, True, upd)
    f_decoder_list = []
    for j in range(cfg.num_layers):
        f_interp_i = nearest_interpolation(feature, inputs["interp_idx"][-j - 1])
        f_decoder_i = conv2d_transpose(torch.cat([f_encoder_list[-j - 2], f_interp_i], dim=3), p,
                                       "Decoder_layer_" + str(j), is_training, upd)
        feature = f_decoder_i
        f_decoder_list.append(f_decoder_i)
    f_layer_fc1 = conv2d(f_decoder_list[-1], p, "fc1", True, is_training, True, upd)
    f_layer_fc2 = conv2d(f_layer_fc1, p, "fc2", True, is_training, True, upd)
    if is_training:
        # tf.nn.dropout(x, keep_prob=0.5): x * mask / keep_prob ; mask injected (shared with the CUDA path)
        f_layer_drop = f_layer_fc2 * dropout_mask.to(f_layer_fc2.dtype) / 0.5 if dropout_mask is not None else f_layer_fc2
    else:
        f_layer_drop = f_layer_fc2
    f_layer_fc3 = conv2d(f_layer_drop, p, "fc", False, is_training, False, upd)
    return f_layer_fc3.squeeze(2)


def get_loss(logits, labels, class_weights):
    """RandLANet.py:62-84,267-274 with no ignored labels: class-weighted softmax CE, mean over points."""
    C = logits.shape[-1]
    logits = logits.reshape(-1, C)
    labels = labels.reshape(-1).long()
    cw = torch.as_tensor(np.asarray(class_weights).reshape(-1), dtype=logits.dtype)
    one_hot = torch.nn.functional.one_hot(labels, C).to(logits.dtype)
    weights = (cw * one_hot).sum(dim=1)
    unweighted = -(one_hot * torch.log_softmax(logits, dim=1)).sum(dim=1)
    return (unweighted * weights).mean()


def tf_map(xyz, cfg, knn):
    """runPancreas.py:124-145: the 5-level index pyramid.  ``knn(support, query, k)`` -> int32 [B,N2,k] numpy."""
    xyz = np.asarray(xyz, dtype=np.float32)
    out = dict(xyz=[], neigh_idx=[], sub_idx=[], interp_idx=[])
    for i in range(cfg.num_layers):
        neigh = knn(xyz, xyz, cfg.k_n)
        n_sub = xyz.shape[1] // cfg.sub_sampling_ratio[i]
        sub_points = xyz[:, :n_sub, :]
        pool_i = neigh[:, :n_sub, :]
        up_i = knn(np.ascontiguousarray(sub_points), xyz, 1)
        out["xyz"].append(xyz)
        out["neigh_idx"].append(neigh)
        out["sub_idx"].append(pool_i)
        out["interp_idx"].append(up_i)
        xyz = np.ascontiguousarray(sub_points)
    return out


def point2prod(list_point_prod, list_xyz, volume_shape, point_idx=None):
    """testPancreas.py:71-85: the per-point Python loop, verbatim semantics (later points overwrite earlier ones),
    followed by np.moveaxis(volume, 1, 2).  ``point_idx`` restates testBraTS.py:226-231 (``test_probs[p_idx] = probs``
    into a zero array over all brain points, then the loop over ALL of them)."""
    list_point_prod = np.asarray(list_point_prod)
    list_xyz = np.asarray(list_xyz)
    if point_idx is not None:
        test_probs = np.zeros((list_xyz.shape[0], list_point_prod.shape[1]), dtype=np.float32)
        test_probs[np.asarray(point_idx)] = list_point_prod
        list_point_prod = test_probs
    volume = np.zeros(volume_shape)
    for i in range(len(list_point_prod)):
        volume[list_xyz[i][2]][list_xyz[i][0]][list_xyz[i][1]] = list_point_prod[i]
    return np.moveaxis(volume, 1, 2)
