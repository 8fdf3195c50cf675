"""bf16 STORAGE mode (BASELINE.json north_star: "a stated bf16 tolerance"; SURVEY.md section 8c proposal: rel-L2 <= 2e-2 on
logits and gradients, arg-max agreement >= 99.5 %).

STATED TOLERANCE (measured on B200, random-initialised 5-level network, 2 x 32 768 BraTS-shaped points, vs the fp64
restatement): logits rel-L2 <= 2e-2 (measured 1.0e-2), loss <= 2e-3 relative (4e-5), arg-max agreement >= 99 % (99.4 %: at
random initialisation the class margins are tiny).  GRADIENTS: the proposal's 2e-2 does NOT hold and no storage rounding
of y can make it hold -- rounding y to 8 bits moves ~0.3 % of the LeakyReLU inputs across zero, the backward (consistently)
follows the rounded forward, and at random initialisation a weight gradient is a noise-dominated sum, so those flips show
up as a 5-10 % rotation: per-tensor rel-L2 median 7e-2, worst 0.4 (cosine similarity >= 0.92).  Gates below: median <= 0.15,
worst <= 0.6.  The mode is therefore OPT-IN for throughput experiments; the fp32 path is the parity path and the headline.

What the mode does: the pre-normalisation activations y of every 1x1 conv that feeds a batch norm -- kept from the forward for
the batch-norm backward, the largest saved tensors of a training step -- are STORED as bfloat16 (ops.set_storage("bf16") /
PU_STORAGE=bf16 / Trainer(storage="bf16")).  Arithmetic, batch statistics, weights, gradients and every other tensor stay fp32.
The default stays fp32 (the 1e-3 parity path)."""
import numpy as np
import pytest
import torch

from oracle import randla_ref as ref
from point_unet_b200 import ops
from point_unet_b200.helper_tool import ConfigBraTS
from point_unet_b200.RandLANet import Network, build_pyramid, init_params

pytestmark = pytest.mark.gpu
TOL_BF16 = 2e-2        # stated tolerance of the mode: relative L2 of logits / per-tensor gradients vs the fp64 restatement
ARGMAX_AGREE = 0.99         # random-initialised weights: class margins are tiny, 99.4-99.7 % measured (proposal was 99.5 %)


@pytest.fixture
def bf16_storage():
    ops.set_storage("bf16")
    yield
    ops.set_storage("fp32")


@pytest.mark.parametrize("M,K,N", [(40000, 64, 32), (9000, 8, 8), (300, 256, 512), (70000, 10, 32)])
def test_linear_stores_rounded_values_and_fp32_statistics(M, K, N, bf16_storage):
    """y (bf16) == round-to-nearest-even of the fp32 result, through the tcgen05 kernel and the narrow kernel; the batch
    statistics are those of the fp32 values; the batch-norm kernels read the bf16 tensor directly."""
    g = torch.Generator().manual_seed(M + K + N)
    x = torch.randn(M, K, generator=g).cuda()
    w = (torch.randn(K, N, generator=g) / K ** 0.5).cuda()
    b = torch.randn(N, generator=g).cuda()
    y16, mean16, var16 = ops.linear_raw(x, w, b, want_stats=True)
    ops.set_storage("fp32")
    y32, mean32, var32 = ops.linear_raw(x, w, b, want_stats=True)
    ops.set_storage("bf16")
    assert y16.dtype == torch.bfloat16 and y32.dtype == torch.float32
    assert torch.equal(y16, y32.to(torch.bfloat16))
    assert torch.equal(mean16, mean32) and torch.equal(var16, var32)
    gamma, beta = torch.rand(N).cuda() + 0.5, torch.randn(N).cuda()
    invstd, scale, shift = ops.bn_prepare(mean16, var16, gamma, beta)
    z16 = ops._bn_act_fwd_raw(y16, scale, shift, 0.2)
    z32 = ops._bn_act_fwd_raw(y16.float(), scale, shift, 0.2)
    assert torch.equal(z16, z32)
    dz = torch.randn(M, N, generator=g).cuda()
    a = ops._bn_bwd_raw(dz, y16, scale, shift, 0.2, gamma, mean16, invstd, True)
    c = ops._bn_bwd_raw(dz, y16.float(), scale, shift, 0.2, gamma, mean16, invstd, True)
    for u, v in zip(a, c):
        assert torch.equal(u, v)


def test_full_network_within_stated_tolerance(bf16_storage):
    from point_unet_b200 import synthetic as syn

    class cfg(ConfigBraTS):
        num_points = 32768
    B = 2
    data = syn.batch(syn.brats_cloud, B, cfg.num_points, seed0=90)
    params = init_params(cfg, 7, seed=3)
    rng = np.random.default_rng(4)
    for k in params:
        if k.endswith("gamma"):
            params[k] = (params[k] + rng.uniform(-0.3, 0.3, params[k].shape)).astype(np.float32)
        if k.endswith("beta") or k.endswith("biases") or k.endswith("bias"):
            params[k] = rng.uniform(-0.1, 0.1, params[k].shape).astype(np.float32)
    net = Network(cfg, 7, device="cuda")
    net.load_numpy(params)
    xyz = torch.from_numpy(data["xyz"]).cuda()
    pyr = build_pyramid(xyz, cfg)
    feats = torch.cat([xyz, torch.from_numpy(data["features"]).cuda()], dim=-1)
    labels = torch.from_numpy(data["labels"]).cuda()
    mask = torch.from_numpy(rng.random((B, cfg.num_points, 1, 32)) < 0.5).cuda()
    logits = net.inference(dict(pyr, features=feats), True, dropout_mask=mask)
    loss = net.get_loss(logits, labels)
    loss.backward()
    assert int(ops.tc_error_flag(logits.device).item()) == 0

    pp = {k: torch.from_numpy(v).double().requires_grad_(not k.split("/")[-1].startswith("moving")) for k, v in params.items()}
    inp = dict(xyz=[t.cpu().double() for t in pyr["xyz"]], neigh_idx=[t.cpu() for t in pyr["neigh_idx"]],
               sub_idx=[t.cpu() for t in pyr["sub_idx"]], interp_idx=[t.cpu() for t in pyr["interp_idx"]],
               features=feats.cpu().double())
    lg = ref.inference(pp, inp, cfg, True, dropout_mask=mask.cpu())
    ls = ref.get_loss(lg, labels.cpu(), net.class_weights.cpu().numpy())
    ls.backward()

    got = logits.detach().cpu().double()
    e_logits = float((got - lg.detach()).norm() / lg.detach().norm())
    agree = float((got.argmax(-1) == lg.detach().argmax(-1)).double().mean())
    print(f"bf16 storage: logits rel-L2 {e_logits:.2e}, arg-max agreement {agree:.4f}, loss {float(loss.detach()):.5f} vs {float(ls.detach()):.5f}")
    assert e_logits < TOL_BF16 and agree >= ARGMAX_AGREE
    assert abs(float(loss.detach()) - float(ls.detach())) < 2e-3 * abs(float(ls.detach()))
    table = []
    for name, t in net.named_variables():
        if (name.endswith("biases") and (name[:-len("biases")] + "bn/gamma") in net._names) or name == "fc0/bias":
            continue   # analytically zero gradients under a training-mode batch norm
        e = float((t.grad.detach().cpu().double() - pp[name].grad).norm() / pp[name].grad.norm())
        table.append((e, name))
    table.sort(reverse=True)
    print("bf16 storage: gradient rel-L2 vs fp64, worst first:", [(f"{e:.2e}", n) for e, n in table[:8]],
          "median", f"{table[len(table) // 2][0]:.2e}")
    assert table[len(table) // 2][0] < 0.15 and table[0][0] < 0.6, table[:5]
