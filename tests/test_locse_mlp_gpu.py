"""GPU parity of the fused position branch of building_block (csrc/locse_mlp.cu) against the oracle restatement in fp64
(oracle/randla_ref.py: relative_pos_encoding -> conv2d 'mlp1' -> BN -> LeakyReLU -> concat with the gathered features,
RandLANet.py:323-328) and against the unfused kernels.  Tolerance 1e-3 relative per tensor (north_star); measured ~1e-6."""
import numpy as np
import pytest
import torch

from oracle import randla_ref as ref
from point_unet_b200 import ops

pytestmark = pytest.mark.gpu
TOL = 1e-3


def rel_err(a, b):
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    assert a.shape == b.shape, (a.shape, b.shape)
    return float((a - b).abs().max() / max(float(b.abs().max()), 1e-12))


def make_case(B, N, K, h, seed, offset=0.0):
    g = torch.Generator().manual_seed(seed)
    xyz = torch.rand(B, N, 3, generator=g) * 2.0 + offset   # an off-centre cloud: the centring of x must cope with it
    idx = torch.randint(0, N, (B, N, K), generator=g, dtype=torch.int32)
    idx[:, :, 0] = torch.arange(N, dtype=torch.int32)[None]
    f_pc = torch.randn(B, N, h, generator=g)
    p = {"s/weights": torch.randn(10, h, generator=g) * (2 / h) ** 0.5, "s/biases": torch.randn(h, generator=g) * 0.1,
         "s/bn/gamma": torch.rand(h, generator=g) + 0.5, "s/bn/beta": torch.randn(h, generator=g) * 0.1,
         "s/bn/moving_mean": torch.randn(h, generator=g) * 0.1, "s/bn/moving_variance": torch.rand(h, generator=g) + 0.5}
    d_buf = torch.randn(B, N, K, 2 * h, generator=g)
    d_fxyz = torch.randn(B, N, K, h, generator=g)
    return xyz, idx, f_pc, p, d_buf, d_fxyz


def oracle_branch(xyz, idx, f_pc, p, training, upd=None):
    f_xyz = ref.relative_pos_encoding(xyz, idx)
    f_xyz = ref.conv2d(f_xyz, p, "s", True, training, True, upd)
    return torch.cat([ref.gather_neighbour(f_pc, idx), f_xyz], dim=-1), f_xyz


@pytest.mark.parametrize("B,N,K,h,offset", [(2, 1500, 16, 8, 0.0), (3, 700, 16, 32, 50.0), (1, 333, 5, 64, 0.0),
                                            (2, 300, 16, 128, 3.0), (1, 130, 16, 256, 0.0), (1, 1, 16, 8, 0.0),
                                            (2, 257, 3, 16, 0.0), (1, 100, 16, 4, 0.0), (4, 9000, 16, 8, 0.0)])
def test_locse_mlp_training_fwd_bwd_vs_oracle(B, N, K, h, offset):
    if not ops.locse_mlp_supported(K, h):
        pytest.skip("fused LocSE branch switched off")
    xyz, idx, f_pc, p, d_buf, d_fxyz = make_case(B, N, K, h, 100 + h, offset)
    pr = {k: v.double().requires_grad_(not k.startswith("s/bn/moving")) for k, v in p.items()}
    fr = f_pc.double().requires_grad_(True)
    upd = {}
    cat_r, fx_r = oracle_branch(xyz.double(), idx, fr, pr, True, upd)
    ((cat_r * d_buf.double()).sum() + (fx_r * d_fxyz.double()).sum()).backward()

    pg = {k: v.cuda().requires_grad_(not k.startswith("s/bn/moving")) for k, v in p.items()}
    fg = f_pc.cuda().requires_grad_(True)
    rows_n = B * N * K
    unbias = rows_n / max(rows_n - 1, 1)
    cat, fx = ops.locse_mlp_concat(xyz.cuda(), fg, idx.cuda(), pg["s/weights"], pg["s/biases"], pg["s/bn/gamma"], pg["s/bn/beta"],
                                   True, pg["s/bn/moving_mean"], pg["s/bn/moving_variance"], unbias, True)
    assert cat.shape == (B, N, K, 2 * h) and fx.shape == (B, N, K, h)
    assert torch.equal(cat[..., :h].cpu(), cat_r[..., :h].detach().float())   # the gathered half is a copy
    assert torch.equal(cat[..., h:], fx)
    assert rel_err(cat, cat_r) < TOL * 0.1
    # moving statistics: momentum 0.99 with TF's unbiased-variance quirk (helper_tf_util.py:553-574)
    mean, var, cnt = upd["s/bn"]
    want_mm = 0.99 * p["s/bn/moving_mean"].double() + 0.01 * mean
    want_mv = 0.99 * p["s/bn/moving_variance"].double() + 0.01 * var * cnt / max(cnt - 1, 1)
    assert rel_err(pg["s/bn/moving_mean"], want_mm) < 1e-5
    assert rel_err(pg["s/bn/moving_variance"], want_mv) < 1e-5
    if rows_n < 64:   # a batch norm over a handful of rows: the backward is all cancellation, forward checked above
        return
    ((cat * d_buf.cuda()).sum() + (fx * d_fxyz.cuda()).sum()).backward()
    assert rel_err(fg.grad, fr.grad) < 1e-5
    for k in ("s/weights", "s/bn/gamma", "s/bn/beta"):
        assert rel_err(pg[k].grad, pr[k].grad) < TOL, (k, rel_err(pg[k].grad, pr[k].grad))
    assert float(pg["s/biases"].grad.abs().max()) == 0.0   # analytically zero through a training-mode batch norm


def test_locse_mlp_matches_unfused_kernels_and_is_deterministic():
    B, N, K, h = 2, 4000, 16, 32
    xyz, idx, f_pc, p, d_buf, d_fxyz = make_case(B, N, K, h, 7)
    xyz, idx, d_buf, d_fxyz = xyz.cuda(), idx.cuda(), d_buf.cuda(), d_fxyz.cuda()

    def run(fused):
        pg = {k: v.cuda().requires_grad_(not k.startswith("s/bn/moving")) for k, v in p.items()}
        fg = f_pc.cuda().requires_grad_(True)
        if fused:
            cat, fx = ops.locse_mlp_concat(xyz, fg, idx, pg["s/weights"], pg["s/biases"], pg["s/bn/gamma"], pg["s/bn/beta"], True,
                                           pg["s/bn/moving_mean"], pg["s/bn/moving_variance"], 1.0, False)
        else:
            x = ops.relative_pos_encoding(xyz, idx)
            y, m, v = ops.linear(x, pg["s/weights"], pg["s/biases"], want_stats=True, zero_bias_grad=True, defer_stats=True)
            cat, fx = ops.lfa_concat(fg, idx, y, m, v, pg["s/bn/gamma"], pg["s/bn/beta"], True, None, need_fxyz=True)
        ((cat * d_buf).sum() + (fx * d_fxyz).sum()).backward()
        return [cat.detach(), fg.grad] + [pg[k].grad for k in ("s/weights", "s/bn/gamma", "s/bn/beta")]

    a, b, c = run(True), run(True), run(False)
    for u, v in zip(a, b):
        assert torch.equal(u, v)            # bit-deterministic
    for u, v in zip(a, c):
        assert rel_err(u, v) < 1e-4


@pytest.mark.parametrize("h", [8, 64])
def test_locse_mlp_inference_mode(h):
    B, N, K = 2, 900, 16
    xyz, idx, f_pc, p, d_buf, d_fxyz = make_case(B, N, K, h, 31 + h)
    pr = {k: v.double().requires_grad_(not k.startswith("s/bn/moving")) for k, v in p.items()}
    fr = f_pc.double().requires_grad_(True)
    cat_r, fx_r = oracle_branch(xyz.double(), idx, fr, pr, False)
    ((cat_r * d_buf.double()).sum() + (fx_r * d_fxyz.double()).sum()).backward()
    pg = {k: v.cuda().requires_grad_(not k.startswith("s/bn/moving")) for k, v in p.items()}
    fg = f_pc.cuda().requires_grad_(True)
    mm0 = pg["s/bn/moving_mean"].clone()
    cat, fx = ops.locse_mlp_concat(xyz.cuda(), fg, idx.cuda(), pg["s/weights"], pg["s/biases"], pg["s/bn/gamma"], pg["s/bn/beta"],
                                   False, pg["s/bn/moving_mean"], pg["s/bn/moving_variance"], 1.0, False)
    assert rel_err(cat, cat_r) < TOL * 0.1
    assert torch.equal(pg["s/bn/moving_mean"], mm0)   # untouched at inference
    ((cat * d_buf.cuda()).sum() + (fx * d_fxyz.cuda()).sum()).backward()
    for k in ("s/weights", "s/biases", "s/bn/gamma", "s/bn/beta"):
        assert rel_err(pg[k].grad, pr[k].grad) < TOL, k
    assert rel_err(fg.grad, fr.grad) < 1e-5


def test_locse_mlp_argument_validation():
    L = ops._L()
    assert L.pu_locse_mlp_supported(16, 8) == 1 and L.pu_locse_mlp_supported(16, 24) == 0
    x = torch.zeros(1, 8, 4, device="cuda")
    i = torch.zeros(1, 8, 16, dtype=torch.int32, device="cuda")
    mom = torch.zeros(65, device="cuda")
    assert L.pu_locse_moments(x.data_ptr(), i.data_ptr(), 1, 8, 16, mom.data_ptr(), None, 0, None) == -2   # PU_ERR_WORKSPACE
    assert L.pu_locse_moments(None, i.data_ptr(), 1, 8, 16, mom.data_ptr(), None, 0, None) == -1            # PU_ERR_INVALID_ARG


@pytest.mark.parametrize("B,N", [(1, 1), (2, 777), (3, 20001)])
def test_att_pool_split_equals_concat(B, N):
    """att16 with the halves of the feature set in two tensors == the same kernels on the concatenated tensor, bit for bit
    (only the addressing differs), forward and backward."""
    K, h = 16, 8
    if not ops.att_pool_split_supported(K, 2 * h):
        pytest.skip("split att16 switched off")
    g = torch.Generator().manual_seed(N)
    left = torch.randn(B, N, K, h, generator=g).cuda()
    right = torch.randn(B, N, K, h, generator=g).cuda()
    w = (torch.randn(2 * h, 2 * h, generator=g) * 0.3).cuda()
    go = torch.randn(B, N, 1, 2 * h, generator=g).cuda()
    l1, r1, w1 = left.clone().requires_grad_(True), right.clone().requires_grad_(True), w.clone().requires_grad_(True)
    o1 = ops.att_pool_split(l1, r1, w1)
    (o1 * go).sum().backward()
    cat = torch.cat([left, right], dim=-1).requires_grad_(True)
    w2 = w.clone().requires_grad_(True)
    o2 = ops.att_pool(cat, w2)
    (o2 * go).sum().backward()
    assert torch.equal(o1, o2)
    assert torch.equal(l1.grad, cat.grad[..., :h]) and torch.equal(r1.grad, cat.grad[..., h:])
    assert torch.equal(w1.grad, w2.grad)
    # and against the oracle in fp64
    pr = {"afc/kernel": w.double().cpu().requires_grad_(True)}
    xr = torch.cat([left, right], dim=-1).double().cpu().requires_grad_(True)
    f = xr.reshape(-1, K, 2 * h)
    s = torch.softmax(f @ pr["afc/kernel"], dim=1)
    want = (f * s).sum(dim=1).reshape(B, N, 1, 2 * h)
    (want * go.double().cpu()).sum().backward()
    assert rel_err(o1, want) < TOL
    assert rel_err(l1.grad, xr.grad[..., :h]) < TOL and rel_err(r1.grad, xr.grad[..., h:]) < TOL
    assert rel_err(w1.grad, pr["afc/kernel"].grad) < TOL


def test_locse_mlp_without_concat_sums_both_consumers():
    """f_pc=None: the result comes back as two aliases; the gradients of both reach the backward kernel and are summed."""
    B, N, K, h = 2, 3000, 16, 8
    xyz, idx, f_pc, p, d_buf, d_fxyz = make_case(B, N, K, h, 55)
    d1, d2 = d_buf[..., :h].contiguous().cuda(), d_fxyz.cuda()

    def run(split):
        pg = {k: v.cuda().requires_grad_(not k.startswith("s/bn/moving")) for k, v in p.items()}
        args = (pg["s/weights"], pg["s/biases"], pg["s/bn/gamma"], pg["s/bn/beta"], True, pg["s/bn/moving_mean"],
                pg["s/bn/moving_variance"], 1.0, False)
        if split:
            a, b = ops.locse_mlp_concat(xyz.cuda(), None, idx.cuda(), *args)
            assert a.data_ptr() == b.data_ptr()
        else:
            cat, b = ops.locse_mlp_concat(xyz.cuda(), f_pc.cuda(), idx.cuda(), *args)
            a = cat[..., h:]
        ((a * d1).sum() + (b * d2).sum()).backward()
        return [b.detach()] + [pg[k].grad for k in ("s/weights", "s/bn/gamma", "s/bn/beta")]

    for u, v in zip(run(True), run(False)):
        assert rel_err(u, v) < 1e-6


@pytest.mark.parametrize("fused,split", [(False, False), (True, False)])
def test_network_switches_agree_with_default_path(fused, split):
    """PU_LOCSE_FUSED / PU_ATT16_SPLIT select older kernel paths (one concat buffer; separate LocSE / conv / BN kernels): the
    whole network must give the same logits and gradients either way (different kernels, same mathematics)."""
    from point_unet_b200 import synthetic as syn
    from point_unet_b200.helper_tool import ConfigBraTS
    from point_unet_b200.RandLANet import Network, build_pyramid

    class Cfg(ConfigBraTS):
        num_points = 8192

    B = 2
    data = syn.batch(syn.brats_cloud, B, Cfg.num_points, seed0=3)
    xyz = torch.from_numpy(data["xyz"]).cuda()
    feats = torch.cat([xyz, torch.from_numpy(data["features"]).cuda()], dim=-1)
    labels = torch.from_numpy(data["labels"]).cuda()
    mask = (torch.rand(B, Cfg.num_points, 1, 32, generator=torch.Generator().manual_seed(1)) < 0.5).cuda()

    def run():
        net = Network(Cfg, feats.shape[-1], seed=5, device="cuda")
        pyr = build_pyramid(xyz, Cfg, locse=True)
        logits = net.inference(dict(pyr, features=feats), True, dropout_mask=mask)
        net.get_loss(logits, labels).backward()
        return logits.detach(), {n: t.grad.clone() for n, t in net.named_variables()}, dict(net.stats)

    keep = (ops.LOCSE_FUSED, ops.ATT16_SPLIT)
    try:
        ref_logits, ref_grads, ref_stats = run()
        ops.LOCSE_FUSED, ops.ATT16_SPLIT = fused, split
        logits, grads, stats = run()
    finally:
        ops.LOCSE_FUSED, ops.ATT16_SPLIT = keep
    assert rel_err(logits, ref_logits) < 1e-4
    for n in ref_stats:
        assert rel_err(stats[n], ref_stats[n]) < 1e-4, n
    bad = []
    for n, g in ref_grads.items():
        scale = float(g.abs().max())
        if scale < 1e-8:      # a bias under a batch norm: analytically zero
            continue
        e = float((grads[n] - g).double().norm() / max(float(g.double().norm()), 1e-30))
        if e > 5e-3:          # routing flips (max-pool winners, LeakyReLU signs) between two fp32 evaluations, see DESIGN
            bad.append((n, e))
    assert not bad, bad
