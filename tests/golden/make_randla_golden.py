"""Generates tests/golden/randla_golden.npz from the fp64 run of the TF-graph restatement (oracle/randla_ref.py) with
indices from the canonical-rule KNN oracle.  python tests/golden/make_randla_golden.py

The fixture holds the inputs (a seeded BraTS-shaped batch), the dropout mask, and the oracle's logits, loss and a
selection of gradients for a 3-level PointSegment (d_out [16,64,128], ratios [4,4,4]) -- deep enough to cover every op
(fc0, dilated_res_block x3 incl. both att_pooling blocks, random_sample, decoder with nearest_interpolation and
conv2d_transpose, head with dropout), small enough (2 x 4096 points) for a 300 KB file.  Variables are regenerated
from `init_params(cfg, 7, seed=5)` plus the perturbation below, so they are not stored."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import knn as ok  # noqa: E402
from oracle import randla_ref as ref  # noqa: E402
from point_unet_b200 import synthetic as syn  # noqa: E402
from point_unet_b200.helper_tool import ConfigBraTS, DataProcessing as DP  # noqa: E402
from point_unet_b200.RandLANet import init_params  # noqa: E402


class GoldenCfg(ConfigBraTS):
    num_points = 4096
    num_layers = 3
    d_out = [16, 64, 128]
    sub_sampling_ratio = [4, 4, 4]


GRAD_KEYS = ["fc0/kernel", "Encoder_layer_0LFAatt_pooling_1fc/kernel", "Encoder_layer_1LFAmlp1/weights",
             "Encoder_layer_1LFAatt_pooling_2fc/kernel", "Encoder_layer_1mlp2/bn/gamma", "Encoder_layer_2shortcut/weights",
             "decoder_0/weights", "Decoder_layer_0/weights", "Decoder_layer_2/bn/beta", "fc1/weights", "fc/weights", "fc/biases"]


def golden_params():
    params = init_params(GoldenCfg, 7, seed=5)
    rng = np.random.default_rng(6)
    for k in params:
        if k.endswith("gamma"):
            params[k] = (params[k] + rng.uniform(-0.3, 0.3, params[k].shape)).astype(np.float32)
        if k.endswith("beta") or k.endswith("biases") or k.endswith("bias"):
            params[k] = rng.uniform(-0.1, 0.1, params[k].shape).astype(np.float32)
    return params


def run_oracle(xyz, feats, labels, mask, dtype=torch.float64):
    cfg = GoldenCfg
    params = golden_params()
    pyr = ref.tf_map(xyz, cfg, lambda s, q, k: ok.knn_restated(s, q, k, tie_rule=1))
    p = {k: torch.from_numpy(v).to(dtype).requires_grad_("moving" not in k) for k, v in params.items()}
    inp = dict(xyz=[torch.from_numpy(a).to(dtype) for a in pyr["xyz"]], neigh_idx=[torch.from_numpy(a) for a in pyr["neigh_idx"]],
               sub_idx=[torch.from_numpy(a) for a in pyr["sub_idx"]], interp_idx=[torch.from_numpy(a) for a in pyr["interp_idx"]],
               features=torch.from_numpy(np.concatenate([xyz, feats], -1)).to(dtype))
    logits = ref.inference(p, inp, cfg, True, dropout_mask=torch.from_numpy(mask))
    loss = ref.get_loss(logits, torch.from_numpy(labels), DP.get_class_weights("BraTS20"))
    loss.backward()
    return pyr, logits.detach(), float(loss.detach()), {k: p[k].grad.detach() for k in GRAD_KEYS}


def main():
    data = syn.batch(syn.brats_cloud, 2, GoldenCfg.num_points, seed0=70)
    mask = np.random.default_rng(8).random((2, GoldenCfg.num_points, 1, 32)) < 0.5
    pyr, logits, loss, grads = run_oracle(data["xyz"], data["features"], data["labels"], mask)
    out = dict(xyz=data["xyz"], features=data["features"], labels=data["labels"].astype(np.int32), mask=np.packbits(mask),
               logits=logits.numpy(), loss=np.float64(loss), neigh_idx_0=pyr["neigh_idx"][0][:, :256], interp_idx_1=pyr["interp_idx"][1])
    for k, g in grads.items():
        out["grad/" + k] = g.numpy().astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "randla_golden.npz"), **out)
    print("loss", loss, "logits", logits.shape, {k: tuple(v.shape) for k, v in grads.items()})


if __name__ == "__main__":
    main()
