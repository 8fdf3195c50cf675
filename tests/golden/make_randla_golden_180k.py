"""Generates tests/golden/randla_golden_180k.npz: the fp64 run of the TF-graph restatement (oracle/randla_ref.py) on the
BENCHMARK configuration -- BASELINE.json configs[2], four BraTS-shaped clouds of 180 000 points (the very clouds bench.py
times: brats_cloud(180000, seed 0..3)), forward + backward -- and configs[0], one Pancreas-shaped 180 000-point cloud,
forward.  Indices come from the canonical-rule KNN oracle.

    python tests/golden/make_randla_golden_180k.py        (about 10 min and 40 GB of host memory on 8 cores)

A 4 x 180k fp64 run materialises ~50 GB of [B,N,K,d] intermediates, so the encoder blocks are recomputed in the backward
(`checkpoint=True`: same arithmetic).  The fixture cannot hold 5 M gradient values per precision, so it stores
  * sha256 digests of the twenty index tensors of the pyramid (the CUDA pyramid must reproduce them bit for bit),
  * every 16th point of the logits, the loss,
  * per trainable variable: the L2 norm of the fp64 gradient, a strided sample of <= 16 384 values, and `e32` = the
    relative L2 deviation of the plain fp32 run of the SAME restatement from the fp64 run (what any fp32 implementation
    of this graph achieves; max-pool / LeakyReLU routing flips under rounding show up here first).
Variables, dropout mask and clouds are regenerated from seeds by `inputs()` below, so they are not stored."""
import hashlib
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import knn as ok  # noqa: E402
from oracle import randla_ref as ref  # noqa: E402
from point_unet_b200 import synthetic as syn  # noqa: E402
from point_unet_b200.helper_tool import ConfigBraTS, ConfigPancreas, DataProcessing as DP  # noqa: E402
from point_unet_b200.RandLANet import init_params  # noqa: E402

OUT = os.path.join(HERE, "randla_golden_180k.npz")
N_POINTS = 180000
LOGIT_STRIDE = 16
GRAD_SAMPLE = 16384


class BenchCfg(ConfigBraTS):
    num_points = N_POINTS


class PancreasCfg(ConfigPancreas):
    num_points = N_POINTS


def perturbed_params(cfg, n_feat, seed):
    """Reference initialisers with gamma / beta / biases moved off their trivial values (their gradients are exercised)."""
    params = init_params(cfg, n_feat, seed=seed)
    rng = np.random.default_rng(seed + 1)
    for k in params:
        if k.endswith("gamma"):
            params[k] = (params[k] + rng.uniform(-0.3, 0.3, params[k].shape)).astype(np.float32)
        if k.endswith("beta") or k.endswith("biases") or k.endswith("bias"):
            params[k] = rng.uniform(-0.1, 0.1, params[k].shape).astype(np.float32)
    return params


def inputs(kind):
    """(cfg, data dict, params, dropout keep-mask) of the two full-size cases; `brats` is bench.py's batch of rank 0."""
    if kind == "brats":
        cfg, gen, B, F = BenchCfg, syn.brats_cloud, 4, 7
    else:
        cfg, gen, B, F = PancreasCfg, syn.pancreas_cloud, 1, 4
    clouds = [gen(N_POINTS, i) for i in range(B)]
    data = {k: np.stack([c[k] for c in clouds]) for k in ("xyz", "features", "labels")}
    params = perturbed_params(cfg, F, seed=11 if kind == "brats" else 13)
    mask = np.random.default_rng(17).random((B, N_POINTS, 1, 32)) < 0.5
    return cfg, data, params, mask


def digest(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def pyramid_digests(pyr) -> list:
    return [digest(np.asarray(pyr[k][i], dtype=np.int32)) for k in ("neigh_idx", "sub_idx", "interp_idx") for i in range(len(pyr[k]))]


def sample(t: torch.Tensor) -> np.ndarray:
    flat = t.detach().reshape(-1)
    stride = max(1, -(-flat.numel() // GRAD_SAMPLE))
    return flat[::stride].to(torch.float32).numpy()


def run(cfg, data, params, mask, pyr, dtype, backward):
    p = {k: torch.from_numpy(v).to(dtype).requires_grad_(backward and "moving" not in k) for k, v in params.items()}
    inp = dict(xyz=[torch.from_numpy(a).to(dtype) for a in pyr["xyz"]], neigh_idx=[torch.from_numpy(a) for a in pyr["neigh_idx"]],
               sub_idx=[torch.from_numpy(a) for a in pyr["sub_idx"]], interp_idx=[torch.from_numpy(a) for a in pyr["interp_idx"]],
               features=torch.from_numpy(np.concatenate([data["xyz"], data["features"]], -1)).to(dtype))
    with torch.set_grad_enabled(backward):
        logits = ref.inference(p, inp, cfg, True, dropout_mask=torch.from_numpy(mask), checkpoint=backward)
        loss = ref.get_loss(logits, torch.from_numpy(data["labels"]), DP.get_class_weights(cfg.name))
    if backward:
        loss.backward()
    return logits.detach(), float(loss.detach()), ({k: v.grad for k, v in p.items() if v.requires_grad} if backward else None)


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    out = {}
    for kind, backward in (("brats", True), ("pancreas", False)):
        t0 = time.time()
        cfg, data, params, mask = inputs(kind)
        pyr = ref.tf_map(data["xyz"], cfg, lambda s, q, k: ok.knn_restated(s, q, k, tie_rule=1))
        print(kind, "pyramid", round(time.time() - t0, 1), "s", flush=True)
        logits, loss, g64 = run(cfg, data, params, mask, pyr, torch.float64, backward)
        print(kind, "fp64 done", round(time.time() - t0, 1), "s loss", loss, flush=True)
        out[kind + "/inputs_sha256"] = np.array([digest(data["xyz"]), digest(data["features"]), digest(data["labels"].astype(np.int32)),
                                                 digest(np.packbits(mask)), digest(np.concatenate([params[k].ravel() for k in sorted(params)]))])
        out[kind + "/pyramid_sha256"] = np.array(pyramid_digests(pyr))
        out[kind + "/logits"] = logits[:, ::LOGIT_STRIDE].to(torch.float32).numpy()
        out[kind + "/logits_absmax"] = np.float64(logits.abs().max())
        out[kind + "/loss"] = np.float64(loss)
        if backward:
            l32, loss32, g32 = run(cfg, data, params, mask, pyr, torch.float32, True)
            print(kind, "fp32 done", round(time.time() - t0, 1), "s loss", loss32, flush=True)
            out[kind + "/logits_e32"] = np.float64((l32.double() - logits).abs().max() / logits.abs().max())
            for k, g in g64.items():
                out[kind + "/grad/" + k] = sample(g)
                out[kind + "/gnorm/" + k] = np.float64(g.norm())
                out[kind + "/e32/" + k] = np.float64((g32[k].double() - g).norm() / g.norm().clamp_min(1e-300))
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT) // 1024, "KB")


if __name__ == "__main__":
    main()
