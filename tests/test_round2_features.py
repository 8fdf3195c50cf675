"""Round-2 rows: TF checkpoint import by variable name (f4), label volume (f2), volume -> cloud construction on the GPU (f3),
torch.library registration (b2), BN buffers / train-eval / learning-rate schedule / error reporting (advisor findings)."""
import os

import numpy as np
import pytest
import torch

from point_unet_b200 import data_prepare as dp
from point_unet_b200 import tf_checkpoint as tfc
from point_unet_b200.helper_tool import ConfigBraTS, ConfigPancreas
from point_unet_b200.RandLANet import Network, tf_variable_name


class SmallBraTS(ConfigBraTS):
    num_points = 4096


# ------------------------------------------------------------------------------------------------ TF checkpoints (CPU)
def test_crc32c_known_answers():
    # RFC 3720 appendix B.4 test vectors
    assert tfc.crc32c(b"123456789") == 0xE3069283
    assert tfc.crc32c(bytes(32)) == 0x8A9136AA
    assert tfc.crc32c(bytes([0xFF] * 32)) == 0x62A8AB43
    assert tfc.crc32c(bytes(range(32))) == 0x46DD794E
    big = bytes(range(256)) * 64                       # >= 4096 bytes: goes through libpointunet_b200's host helper
    # chaining a pure-Python prefix (< 4096 bytes) into the helper equals the one-shot helper result
    assert tfc.crc32c(big) == tfc.crc32c(big[4000:], tfc.crc32c(big[:4000]))


def test_tf_checkpoint_round_trip_and_table_layout(tmp_path):
    rng = np.random.default_rng(0)
    tensors = {"layers/fc0/kernel": rng.standard_normal((7, 8)).astype(np.float32),
               "layers/batch_normalization/moving_variance": rng.random(8).astype(np.float32),
               "optimizer/beta1_power": np.float32(0.9), "global_step": np.int64(1234)}
    for i in range(300):   # enough keys for several 4 KB index blocks (prefix compression + restarts + index block)
        tensors["layers/Encoder_layer_%dLFAmlp%d/weights" % (i % 5, i)] = rng.standard_normal((1, 1, 3, 5)).astype(np.float32)
    prefix = str(tmp_path / "snapshots" / "snap-1500")
    tfc.write_checkpoint(prefix, tensors)
    assert os.path.exists(prefix + ".index") and os.path.exists(prefix + ".data-00000-of-00001")
    raw = open(prefix + ".index", "rb").read()
    assert raw[-8:] == (0xDB4775248B80FB57).to_bytes(8, "little")     # LevelDB table magic
    back = tfc.read_checkpoint(prefix)
    assert set(back) == set(tensors)
    for k, v in tensors.items():
        assert back[k].dtype == np.asarray(v).dtype and np.array_equal(back[k], v), k
    listing = tfc.list_variables(prefix)
    assert listing["layers/fc0/kernel"] == (np.float32, (7, 8)) and listing["global_step"][1] == ()
    # a flipped data byte is caught by the per-tensor CRC32C, a flipped index byte by the block CRC
    data = bytearray(open(prefix + ".data-00000-of-00001", "rb").read())
    data[10] ^= 0x40
    open(prefix + ".data-00000-of-00001", "wb").write(bytes(data))
    with pytest.raises(ValueError, match="checksum"):
        tfc.read_checkpoint(prefix)
    bad = bytearray(raw)
    bad[20] ^= 0x01
    open(prefix + ".index", "wb").write(bytes(bad))
    with pytest.raises(ValueError, match="checksum"):
        tfc.read_checkpoint(prefix)


def test_tf_variable_names_follow_the_reference_graph():
    # RandLANet.py:56 'layers' scope; helper_tf_util.py:148,167 conv2d scope + unnamed tf.layers.batch_normalization;
    # RandLANet.py:114-115 fc0 dense + the BN directly under 'layers'
    assert tf_variable_name("fc0/kernel") == "layers/fc0/kernel"
    assert tf_variable_name("fc0/bn/gamma") == "layers/batch_normalization/gamma"
    assert tf_variable_name("Encoder_layer_2LFAatt_pooling_1fc/kernel") == "layers/Encoder_layer_2LFAatt_pooling_1fc/kernel"
    assert tf_variable_name("Decoder_layer_3/bn/moving_variance") == "layers/Decoder_layer_3/batch_normalization/moving_variance"
    assert tf_variable_name("fc/biases") == "layers/fc/biases"


def test_network_tf_checkpoint_and_state_dict_round_trip(tmp_path):
    """save -> load by TF variable name restores every variable AND the BN moving statistics (the reference's Saver holds
    all GLOBAL_VARIABLES, RandLANet.py:101); optimizer slots in the snapshot are ignored; state_dict covers the same set."""
    a = Network(SmallBraTS, 7, seed=1, device="cpu")
    rng = np.random.default_rng(3)
    with torch.no_grad():
        for k, t in a.stats.items():
            t.copy_(torch.from_numpy(rng.random(t.shape).astype(np.float32)))
    prefix = str(tmp_path / "snap-7")
    a.save_tf_checkpoint(prefix)
    ck = tfc.read_checkpoint(prefix)
    assert ck["layers/Encoder_layer_0mlp1/weights"].shape == (1, 1, 8, 8)            # helper_tf_util.py:151 4-D 1x1 kernels
    assert ck["layers/Decoder_layer_0/weights"].shape == (1, 1, 512, 1536)           # :211-212 [1,1,Cout,Cin]
    extra = dict(ck)
    extra["layers/fc0/kernel/Adam"] = np.zeros((7, 8), np.float32)                    # optimizer slots are skipped
    extra["optimizer/beta1_power"] = np.float32(0.5)
    tfc.write_checkpoint(prefix, extra)
    b = Network(SmallBraTS, 7, seed=2, device="cpu")
    loaded = b.load_tf_checkpoint(prefix)
    assert len(loaded) == len(a._names) + len(a._stat_names)
    for (n, ta), (_, tb) in zip(a.named_variables(), b.named_variables()):
        assert torch.equal(ta, tb), n
    for k in a.stats:
        assert torch.equal(a.stats[k], b.stats[k]), k
    # a snapshot that lacks a variable is an error unless strict=False
    del extra["layers/fc/weights"]
    tfc.write_checkpoint(prefix, extra)
    with pytest.raises(KeyError):
        Network(SmallBraTS, 7, device="cpu").load_tf_checkpoint(prefix)
    # torch-native persistence: moving statistics are buffers
    sd = a.state_dict()
    assert sum(k.startswith("stat__") for k in sd) == len(a._stat_names) and "class_weights" not in sd
    c = Network(SmallBraTS, 7, seed=5, device="cpu")
    c.load_state_dict(sd)
    for k in a.stats:
        assert torch.equal(a.stats[k], c.stats[k]), k
    # train() / eval() drive the reference's is_training switch
    assert c.is_training and not c.eval().is_training and c.train().is_training


# ------------------------------------------------------------------------------------------------ torch.library (CPU part)
def test_custom_ops_are_registered_with_shape_functions():
    from point_unet_b200 import torch_ops  # noqa: F401
    ns = torch.ops.pointunet
    for name in torch_ops.OP_NAMES:
        assert hasattr(ns, name), name
    m = "meta"
    pc, idx = torch.empty(2, 50, 8, device=m), torch.empty(2, 50, 16, dtype=torch.int32, device=m)
    assert ns.gather_neighbour(pc, idx).shape == (2, 50, 16, 8)
    assert ns.relative_pos_encoding(torch.empty(2, 50, 3, device=m), idx).shape == (2, 50, 16, 10)
    assert ns.att_pooling(torch.empty(2, 50, 16, 32, device=m), torch.empty(32, 32, device=m)).shape == (2, 50, 1, 32)
    assert ns.nearest_interpolation(torch.empty(2, 12, 1, 8, device=m), torch.empty(2, 50, 1, dtype=torch.int32, device=m)).shape == (2, 50, 1, 8)
    assert ns.knn_search(torch.empty(2, 50, 3, device=m), torch.empty(2, 20, 3, device=m), 16).shape == (2, 20, 16)
    out, ties = ns.random_sample_fwd(torch.empty(2, 50, 1, 8, device=m), torch.empty(2, 12, 16, dtype=torch.int32, device=m))
    assert out.shape == (2, 12, 1, 8) and ties.dtype == torch.uint8
    with pytest.raises((NotImplementedError, RuntimeError)):   # no CPU kernel is registered: the product path is CUDA only
        ns.gather_neighbour(torch.zeros(1, 4, 4), torch.zeros(1, 4, 2, dtype=torch.int32))


# ------------------------------------------------------------------------------------------------ data preparation vs oracle
def _volumes(seed=0):
    rng = np.random.default_rng(seed)
    shape = (14, 11, 9)
    img = rng.normal(40.0, 12.0, shape)
    label = np.zeros(shape, np.uint8)
    label[4:8, 3:7, 2:6] = 1
    mods = rng.normal(100.0, 30.0, (4,) + shape)
    brain = np.zeros(shape, bool)
    brain[2:12, 1:10, 1:8] = True
    mods[:, ~brain] = 0.0
    blabel = np.zeros(shape, np.uint8)
    blabel[5:9, 4:8, 3:6] = rng.integers(1, 4, (4, 4, 3))
    return img, label, mods, blabel


def _check_prepare(device):
    from oracle import prepare_ref as pr
    img, label, mods, blabel = _volumes()
    rng = np.random.default_rng(1)
    # Pancreas: z-score over the whole volume, foreground first, injected background draw
    z = dp.zscore_volume(torch.from_numpy(img).to(device), nonzero_only=False)
    assert np.allclose(z.cpu().numpy(), pr.intensity_normalize_whole(img), rtol=1e-12, atol=1e-12)
    n_point = 400
    n_fg = int((label > 0).sum())
    choice = rng.permutation(int((label == 0).sum()))[: n_point - n_fg]
    want = pr.pancreas_cloud(pr.intensity_normalize_whole(img), label, n_point, choice)
    got = dp.sample_pancreas_cloud(z, torch.from_numpy(label).to(device), n_point, background_choice=choice)
    assert np.array_equal(got["xyz_origin"].cpu().numpy(), want["xyz_origin"].astype(np.int32))
    assert got["xyz"].dtype == torch.float32 and np.array_equal(got["xyz"].cpu().numpy(), want["xyz"])  # fp32 division, bit-exact
    assert np.array_equal(got["value"].cpu().numpy(), want["value"]) and np.array_equal(got["labels"].cpu().numpy(), want["labels"])
    # BraTS: per-modality z-score over v > 0, brain = any modality non-zero, fp64 division then cast, shuffle
    zm = torch.stack([dp.zscore_volume(torch.from_numpy(m).to(device), nonzero_only=True) for m in mods])
    zr = np.stack([pr.intensity_normalize_nonzero(m) for m in mods])
    assert np.allclose(zm.cpu().numpy(), zr, rtol=1e-12, atol=1e-12)
    full = pr.brats_full_cloud(np.concatenate([zr, blabel[None].astype(np.float64)]))
    n_pts = 300
    n_t = int((full["labels"] > 0).sum())
    choice = rng.permutation(int((full["labels"] == 0).sum()))[: n_pts - n_t]
    perm = rng.permutation(n_pts)
    want = pr.brats_sample(full, n_pts, choice, perm)
    got = dp.sample_brats_cloud(torch.from_numpy(zr).to(device), torch.from_numpy(blabel).to(device), n_pts,
                                background_choice=choice, shuffle_perm=perm)
    assert np.array_equal(got["xyz_origin_all"].cpu().numpy(), full["xyz_origin"].astype(np.int32))
    assert np.array_equal(got["point_idx"].cpu().numpy(), want["point_idx"])
    assert np.array_equal(got["xyz"].cpu().numpy(), want["xyz"])
    assert np.array_equal(got["colors"].cpu().numpy(), want["colors"]) and np.array_equal(got["labels"].cpu().numpy(), want["labels"])
    return got


def test_cloud_construction_matches_reference_restatement_cpu(tmp_path):
    got = _check_prepare("cpu")
    # <ID>_xyz_origin.npy flavours (dataPreparePancreas.py:160-161 uint16; dataPrepareBraTS.py:81-82 int)
    p = str(tmp_path / "case_xyz_origin.npy")
    dp.save_xyz_origin(p, got["xyz_origin_all"], "BraTS")
    assert np.load(p).dtype == np.dtype(int) and torch.equal(dp.load_xyz_origin(p), got["xyz_origin_all"])
    dp.save_xyz_origin(p, got["xyz_origin_all"], "Pancreas")
    assert np.load(p).dtype == np.uint16 and torch.equal(dp.load_xyz_origin(p), got["xyz_origin_all"])


@pytest.mark.gpu
def test_cloud_construction_matches_reference_restatement_gpu():
    got = _check_prepare("cuda")
    assert got["xyz"].is_cuda


# ------------------------------------------------------------------------------------------------ GPU rows
@pytest.mark.gpu
def test_point2label_vs_reference_argmax():
    """utils/genSegmentationPancreas.py:67-77: seg = argmax(volume, -1).astype(uint8); BraTS: seg[seg == 3] = 4."""
    from oracle import randla_ref as ref
    from point_unet_b200 import ops
    rng = np.random.default_rng(0)
    Z, X, Y, C = 12, 20, 16, 4
    n = 3000
    xyz_o = np.stack([rng.integers(0, X, n), rng.integers(0, Y, n), rng.integers(0, Z, n)], axis=1).astype(np.int32)
    probs = rng.random((n, C), dtype=np.float32)
    probs[::7] = probs[::7, :1]                      # exact ties across classes: the first maximum must win
    vol = ref.point2prod(probs, xyz_o, (Z, X, Y, C))
    want = np.argmax(vol, axis=-1).astype(np.uint8)
    got = ops.point2label(torch.from_numpy(probs).cuda(), torch.from_numpy(xyz_o).cuda(), (Z, X, Y, C))
    assert got.dtype == torch.uint8 and np.array_equal(got.cpu().numpy(), want)
    want4 = want.copy()
    want4[want4 == 3] = 4
    got4 = ops.point2label(torch.from_numpy(probs).cuda(), torch.from_numpy(xyz_o).cuda(), (Z, X, Y, C), remap=(3, 4))
    assert np.array_equal(got4.cpu().numpy(), want4)
    dense = ops.point2prod(torch.from_numpy(probs).cuda(), torch.from_numpy(xyz_o).cuda(), (Z, X, Y, C))
    assert np.array_equal(ops.volume_argmax(dense, remap=(3, 4)).cpu().numpy(), want4)
    # BraTS p_idx expansion
    all_vox = np.stack(np.unravel_index(rng.choice(X * Y * Z, 2000, replace=False), (X, Y, Z)), axis=1).astype(np.int32)
    sel = rng.permutation(2000)[:700].astype(np.int32)
    pr = rng.random((700, C), dtype=np.float32)
    want = np.argmax(ref.point2prod(pr, all_vox, (Z, X, Y, C), point_idx=sel), axis=-1).astype(np.uint8)
    got = ops.point2label(torch.from_numpy(pr).cuda(), torch.from_numpy(all_vox).cuda(), (Z, X, Y, C), point_idx=torch.from_numpy(sel).cuda())
    assert np.array_equal(got.cpu().numpy(), want)


@pytest.mark.gpu
def test_registered_ops_match_the_autograd_shims():
    from point_unet_b200 import ops, torch_ops
    from point_unet_b200.helper_tool import knn_search_cuda
    g = torch.Generator().manual_seed(0)
    B, N, K, d = 2, 3000, 16, 32
    xyz = torch.rand(B, N, 3, generator=g).cuda()
    idx = torch.ops.pointunet.knn_search(xyz, xyz, K)
    assert torch.equal(idx, knn_search_cuda(xyz, xyz, K))
    assert torch.equal(torch.ops.pointunet.relative_pos_encoding(xyz, idx), ops.relative_pos_encoding(xyz, idx))
    pc = torch.randn(B, N, d, generator=g).cuda()
    w = (torch.randn(2 * d, 2 * d, generator=g) * 0.2).cuda()
    pool_idx = idx[:, : N // 4].contiguous()
    interp = knn_search_cuda(xyz[:, : N // 4].contiguous(), xyz, 1)

    def run(f_gather, f_att, f_pool, f_interp):
        p, ww = pc.clone().requires_grad_(True), w.clone().requires_grad_(True)
        nb = f_gather(p, idx)                                            # [B,N,K,d]
        agg = f_att(torch.cat([nb, nb * 0.5], dim=-1), ww)                # [B,N,1,2d]
        pooled = f_pool(agg, pool_idx)                                   # [B,N/4,1,2d]
        up = f_interp(pooled, interp)                                    # [B,N,1,2d]
        (up * up).sum().backward()
        ops.clear_caches()
        return up.detach(), p.grad, ww.grad
    a = run(ops.gather_neighbour, ops.att_pool, ops.random_sample, ops.nearest_interpolation)
    ns = torch.ops.pointunet
    b = run(ns.gather_neighbour, ns.att_pooling, torch_ops.random_sample, ns.nearest_interpolation)
    for x, y in zip(a, b):
        assert torch.equal(x, y)


@pytest.mark.gpu
def test_lr_decay_reaches_captured_graph_and_errors_are_reported():
    """RandLANet.py:190-193 lr *= 0.95 per epoch: the learning rate is a device scalar, so a replayed CUDA graph follows it
    (lr -> 0 freezes the weights); a set tcgen05 error flag makes the host-facing step raise and is then cleared."""
    from point_unet_b200 import _lib, ops, synthetic
    from point_unet_b200.train import Trainer
    data = synthetic.batch(synthetic.brats_cloud, 2, SmallBraTS.num_points, 5)
    x, f, l = (torch.from_numpy(data[k]).cuda() for k in ("xyz", "features", "labels"))
    tr = Trainer(SmallBraTS, num_features=7, seed=0, device="cuda")
    assert abs(float(tr.lr) - 1e-4) < 1e-10
    tr.decay_lr()
    assert abs(float(tr.lr) - 0.95e-4) < 1e-10                          # cfg.lr_decays = 0.95
    tr.capture_step(x, f, l, warmup=1)
    w0 = torch.cat([p.detach().flatten() for p in tr.params]).clone()
    tr.train_step_graph(x, f, l)
    w1 = torch.cat([p.detach().flatten() for p in tr.params]).clone()
    assert not torch.equal(w0, w1)
    tr.decay_lr(0.0)                                                     # lr = 0 -> Adam's update is exactly zero
    tr.train_step_graph(x, f, l)
    w2 = torch.cat([p.detach().flatten() for p in tr.params])
    assert torch.equal(w1, w2)
    # host-facing steps read (loss, error flag) back together
    loss = tr.train_step(data["xyz"], data["features"], data["labels"])
    assert np.isfinite(loss)
    ops.tc_error_flag(torch.device("cuda", torch.cuda.current_device())).fill_(1)
    with pytest.raises(_lib.PointUnetError, match="barrier timed out"):
        tr.train_step(data["xyz"], data["features"], data["labels"])
    assert int(ops.tc_error_flag(torch.device("cuda", torch.cuda.current_device())).item()) == 0
    assert np.isfinite(tr.train_step(data["xyz"], data["features"], data["labels"]))


@pytest.mark.gpu
def test_state_dict_round_trip_keeps_predictions():
    from point_unet_b200 import synthetic
    from point_unet_b200.train import Trainer
    data = synthetic.batch(synthetic.brats_cloud, 2, SmallBraTS.num_points, 9)
    a = Trainer(SmallBraTS, num_features=7, seed=0, device="cuda")
    for _ in range(3):   # moves the BN moving statistics away from (0, 1)
        a.train_step(data["xyz"], data["features"], data["labels"])
    want = a.predict(data["xyz"], data["features"]).clone()
    b = Trainer(SmallBraTS, num_features=7, seed=4, device="cuda")
    b.net.load_state_dict(a.net.state_dict())
    assert torch.equal(b.predict(data["xyz"], data["features"]), want)
    # eval() switches the module's forward to moving statistics, no dropout (deterministic)
    x = torch.from_numpy(data["xyz"]).cuda()
    from point_unet_b200.RandLANet import build_pyramid
    inputs = dict(build_pyramid(x, SmallBraTS), features=torch.cat([x, torch.from_numpy(data["features"]).cuda()], -1))
    b.net.eval()
    with torch.no_grad():
        assert torch.equal(torch.softmax(b.net(inputs), -1), want)


@pytest.mark.gpu
def test_overlapped_host_staging_trains_the_submitted_batches():
    """Trainer.train_step on a captured graph: call k copies batch k on the copy stream while the replay trains the batch
    that landed during call k-1.  The sequence of trained batches (and with a fixed dropout mask the resulting weights) equals
    plain sequential training, one call late."""
    from point_unet_b200 import synthetic
    from point_unet_b200.train import Trainer
    batches = [synthetic.batch(synthetic.brats_cloud, 2, SmallBraTS.num_points, s) for s in (31, 32, 33)]
    dev = [tuple(torch.from_numpy(b[k]).cuda() for k in ("xyz", "features", "labels")) for b in batches]
    mask = torch.rand(2, SmallBraTS.num_points, 1, 32, device="cuda", generator=torch.Generator("cuda").manual_seed(1)) < 0.5
    seq = Trainer(SmallBraTS, num_features=7, seed=0, device="cuda")
    ovl = Trainer(SmallBraTS, num_features=7, seed=0, device="cuda")
    seq.dropout_mask = ovl.dropout_mask = mask
    ls = [float(seq.train_step_device(*dev[i])) for i in (0, 0, 1, 2)]   # warm-up on batch 0, then 0, 1, 2
    ovl.capture_step(*dev[0], warmup=1)
    lo = [ovl.train_step(batches[i]["xyz"], batches[i]["features"], batches[i]["labels"]) for i in (1, 2, 2, 2)]
    # call 0 trains the capture batch (0), call 1 trains batch 1, call 2 trains batch 2 (call 3 trains batch 2 again)
    assert lo[:3] == pytest.approx(ls[1:], rel=0, abs=0), (ls, lo)


@pytest.mark.gpu
@pytest.mark.parametrize("P", [8, 1000, 45001])
def test_fused_att_pooling_backward_d64(P):
    """d = 64: dx = g s + d_act w^T from ONE kernel (second tcgen05 MMA inside the epilogue) against the fp64 restatement
    (RandLANet.py:394-398) and against the two-kernel path (att backward + accumulate GEMM); bit-deterministic."""
    from point_unet_b200 import ops
    K, d = 16, 64
    g = torch.Generator().manual_seed(P)
    x = torch.randn(1, P, K, d, generator=g)
    w = torch.randn(d, d, generator=g) * 0.2
    dy = torch.randn(1, P, 1, d, generator=g)
    xr, wr = x.double().requires_grad_(True), w.double().requires_grad_(True)
    act = xr.reshape(-1, K, d) @ wr
    agg_r = (xr.reshape(-1, K, d) * torch.softmax(act, dim=1)).sum(1).reshape(1, P, 1, d)
    (agg_r * dy.double()).sum().backward()

    def run(fused):
        old = ops.ATT_BWD_FUSED
        ops.ATT_BWD_FUSED = fused
        try:
            xg, wg = x.cuda().requires_grad_(True), w.cuda().requires_grad_(True)
            ops.tc_error_flag(xg.device).zero_()
            agg = ops.att_pool(xg, wg)
            (agg * dy.cuda()).sum().backward()
            assert int(ops.tc_error_flag(xg.device).item()) == 0, "tcgen05 pipeline barrier timed out"
            return xg.grad, wg.grad
        finally:
            ops.ATT_BWD_FUSED = old

    def rel(a, b):
        return float((a.detach().cpu().double() - b).abs().max() / b.abs().max())
    dx1, dw1 = run(True)
    assert rel(dx1, xr.grad) < 1e-4 and rel(dw1, wr.grad) < 1e-4, (rel(dx1, xr.grad), rel(dw1, wr.grad))
    dx0, dw0 = run(False)
    assert rel(dx1, dx0.cpu().double()) < 1e-5 and torch.equal(dw1, dw0)   # d_act (hence dw) is produced identically
    dx2, dw2 = run(True)
    assert torch.equal(dx1, dx2) and torch.equal(dw1, dw2)
