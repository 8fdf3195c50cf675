"""CPU tests of the TF-graph restatement (oracle/randla_ref.py) and of host-side logic shared with the CUDA path."""
import numpy as np
import torch

from oracle import knn as ok
from oracle import randla_ref as ref
from point_unet_b200.helper_tool import ConfigPancreas
from point_unet_b200.ops import rows
from point_unet_b200.RandLANet import init_params, layer_table


def test_att_pooling_matches_loop_definition():
    torch.manual_seed(0)
    B, N, K, d = 1, 5, 16, 8
    x = torch.randn(B, N, K, d, dtype=torch.float64)
    W = torch.randn(d, d, dtype=torch.float64)
    p = {"afc/kernel": W, "amlp/weights": torch.eye(d, dtype=torch.float64), "amlp/biases": torch.zeros(d, dtype=torch.float64),
         "amlp/bn/gamma": torch.ones(d, dtype=torch.float64), "amlp/bn/beta": torch.zeros(d, dtype=torch.float64)}
    keep = {}
    ref.att_pooling(x, p, "a", True, None, keep)
    want = torch.zeros(B, N, d, dtype=torch.float64)
    for n in range(N):
        act = x[0, n] @ W                      # [K, d]
        for c in range(d):
            s = torch.exp(act[:, c] - act[:, c].max())
            s = s / s.sum()                    # softmax over the K neighbours, per channel
            want[0, n, c] = (x[0, n, :, c] * s).sum()
    assert torch.allclose(keep["af_agg"].squeeze(2), want, atol=1e-12)


def test_locse_layout_and_self_distance():
    xyz = torch.rand(2, 50, 3, dtype=torch.float64)
    idx = torch.randint(0, 50, (2, 50, 16), dtype=torch.int32)
    idx[:, :, 0] = torch.arange(50, dtype=torch.int32)
    f = ref.relative_pos_encoding(xyz, idx)
    assert f.shape == (2, 50, 16, 10)
    assert (f[:, :, 0, 0] == 0).all()
    nb = ref.gather_neighbour(xyz, idx)
    assert torch.equal(f[..., 7:10], nb) and torch.equal(f[..., 1:4], f[..., 4:7] - nb)
    assert torch.allclose(f[..., 0], (f[..., 1:4] ** 2).sum(-1).sqrt())


def test_random_sample_gradient_splits_ties_evenly():
    feat = torch.zeros(1, 4, 1, 1, dtype=torch.float64, requires_grad=True)
    idx = torch.tensor([[[0, 1, 2, 3]]], dtype=torch.int32)
    ref.random_sample(feat, idx).sum().backward()
    assert torch.allclose(feat.grad.flatten(), torch.full((4,), 0.25, dtype=torch.float64))


def test_layer_table_and_param_count():
    cfg = ConfigPancreas
    t = layer_table(cfg, 4)
    names = [r[0] for r in t]
    assert names[0] == "fc0" and names[-1] == "fc" and "Encoder_layer_4LFAatt_pooling_2fc" in names
    dec = [r for r in t if r[1] == "convT"]
    assert [(r[2], r[3]) for r in dec] == [(1536, 512), (768, 256), (384, 128), (160, 32), (64, 32)]
    p = init_params(cfg, 4, seed=0)
    n_train = sum(v.size for k, v in p.items() if "moving" not in k)
    assert 4.9e6 < n_train < 5.1e6  # ~4.99 M parameters
    w = p["Encoder_layer_1mlp1/weights"]
    assert np.allclose(w, np.round(w * 1000) / 1000) and abs(w).max() <= 2 * np.sqrt(2 / w.shape[1]) + 1e-3


def test_pyramid_restatement_shapes_with_oracle_knn():
    class Cfg(ConfigPancreas):
        num_points = 2048
    xyz = np.random.default_rng(0).random((2, 2048, 3), dtype=np.float32)
    pyr = ref.tf_map(xyz, Cfg, lambda s, q, k: ok.knn_restated(s, q, k, tie_rule=1))
    assert [a.shape[1] for a in pyr["xyz"]] == [2048, 512, 128, 32, 8]
    assert pyr["sub_idx"][0].shape == (2, 512, 16) and pyr["interp_idx"][0].shape == (2, 2048, 1)
    assert np.array_equal(pyr["sub_idx"][1], pyr["neigh_idx"][1][:, :128])
    assert pyr["interp_idx"][0].max() < 512


def test_small_network_runs_and_grads_flow_fp64():
    class Cfg(ConfigPancreas):
        num_points = 1024
        num_layers = 2
        d_out = [16, 64]
        sub_sampling_ratio = [4, 4]
    rng = np.random.default_rng(1)
    xyz = rng.random((1, 1024, 3), dtype=np.float32)
    pyr = ref.tf_map(xyz, Cfg, lambda s, q, k: ok.knn_restated(s, q, k, tie_rule=1))
    p = {k: torch.from_numpy(v).double().requires_grad_("moving" not in k) for k, v in init_params(Cfg, 4, 0).items()}
    inputs = dict(xyz=[torch.from_numpy(a).double() for a in pyr["xyz"]], neigh_idx=[torch.from_numpy(a) for a in pyr["neigh_idx"]],
                  sub_idx=[torch.from_numpy(a) for a in pyr["sub_idx"]], interp_idx=[torch.from_numpy(a) for a in pyr["interp_idx"]],
                  features=torch.from_numpy(np.concatenate([xyz, rng.standard_normal((1, 1024, 1))], -1)).double())
    logits = ref.inference(p, inputs, Cfg, True, dropout_mask=torch.ones(1, 1024, 1, 32, dtype=torch.bool))
    assert logits.shape == (1, 1024, 2)
    loss = ref.get_loss(logits, torch.from_numpy(rng.integers(0, 2, (1, 1024))), np.array([[1.923, 1.923]]))
    loss.backward()
    assert all(v.grad is not None and torch.isfinite(v.grad).all() for k, v in p.items() if v.requires_grad)
    assert np.isfinite(float(loss)) and float(loss) > 0


def test_rows_view_of_concat_halves():
    buf = torch.zeros(2, 5, 16, 32)
    t, R, C, ld = rows(buf[..., :16])
    assert (R, C, ld) == (160, 16, 32) and t.data_ptr() == buf.data_ptr()
    t, R, C, ld = rows(buf[..., 16:])
    assert (R, C, ld) == (160, 16, 32) and t.data_ptr() == buf[..., 16:].data_ptr()
    t, R, C, ld = rows(buf.transpose(1, 2))  # not expressible as constant-stride rows -> copied
    assert ld == 32 and t.is_contiguous()
    t, R, C, ld = rows(torch.zeros(7, 1, 8).squeeze(1))
    assert (R, C, ld) == (7, 8, 8)


def test_oracle_sources_are_frozen():
    """The restatements are the single point of failure of the TF half ("parity unpinned"): they are frozen.  A deliberate
    edit updates the digest here together with the golden fixtures that were generated from them."""
    import hashlib
    import os
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle")
    want = {"randla_ref.py": "0f73b0757a37a218bf8ea2abaa48b6b01ac077ccae3bd6764a4ccf9bd0b3a9af",
            "knn_oracle.c": "7cd921ae4f9024103ee93f62cb34074188177626edb5e0d9bb43818b4ff63220",
            "prepare_ref.py": "0a596fa462389fc50ee05bc3952f70e4e8bce64472e82cb6040307741af51321"}
    for name, digest in want.items():
        with open(os.path.join(root, name), "rb") as f:
            assert hashlib.sha256(f.read()).hexdigest() == digest, name
