"""Host-side mirror of ``PointSegment/RandLANet.py`` for the hot path: the ``Network`` of RandLA-Net local
feature aggregation, built on the sm_100a kernels in ``csrc/`` (see ``ops.py``).

Same method names, argument order and tensor layouts as the reference (``[B,N,(K,)d]`` channels-last, the
singleton axis-2 convention, int32 indices):

    Network.gather_neighbour(pc, neighbor_idx)            RandLANet.py:377-386
    Network.relative_pos_encoding(xyz, neigh_idx)         RandLANet.py:337-343
    Network.att_pooling(feature_set, d_out, name, is_training)   RandLANet.py:388-401
    Network.random_sample(feature, pool_idx)              RandLANet.py:345-360
    Network.nearest_interpolation(feature, interp_idx)    RandLANet.py:362-375
    Network.building_block / dilated_res_block / inference / get_loss   RandLANet.py:314-335, 110-152, 267-274

Variables keep the reference's scope names (``Encoder_layer_0LFAatt_pooling_1fc/kernel`` ...), stored in the
layouts of ``helper_tf_util.py``: conv2d kernels ``[Cin, Cout]`` (the 1x1 ``[1,1,Cin,Cout]`` squeezed),
conv2d_transpose kernels ``[Cout, Cin]``, dense kernels ``[in, out]``.  Batch-norm variables are named
``<scope>/bn/{gamma,beta,moving_mean,moving_variance}`` here; TensorFlow names them
``layers/<scope>/batch_normalization/...`` (and ``layers/batch_normalization/...`` for the unnamed BN after ``fc0``,
RandLANet.py:115): ``tf_variable_name`` / ``Network.load_tf_checkpoint`` translate.  ``tf_map``
(``runPancreas.py:124-145``) is ``build_pyramid`` here and runs the KNN kernel on the device.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import ops
from .helper_tool import DataProcessing as DP
from .helper_tool import knn_search_cuda, knn_self_interp_cuda

import os as _os

# neigh_idx and interp_idx of a pyramid level from ONE search structure (pu_knn_self_interp); 0 = two separate searches
KNN_FUSED_INTERP = int(_os.environ.get("PU_KNN_FUSED_INTERP", "1")) != 0


def layer_table(cfg, num_features: int):
    """[(scope, kind, Cin, Cout)] in graph order; kind in {dense, conv, convT, att_fc}."""
    t = [("fc0", "dense", num_features, 8)]
    d_in = 8
    for i in range(cfg.num_layers):
        d = cfg.d_out[i]
        n = "Encoder_layer_%d" % i
        t += [(n + "mlp1", "conv", d_in, d // 2),
              (n + "LFAmlp1", "conv", 10, d // 2),
              (n + "LFAatt_pooling_1fc", "att_fc", d, d),
              (n + "LFAatt_pooling_1mlp", "conv", d, d // 2),
              (n + "LFAmlp2", "conv", d // 2, d // 2),
              (n + "LFAatt_pooling_2fc", "att_fc", d, d),
              (n + "LFAatt_pooling_2mlp", "conv", d, d),
              (n + "mlp2", "conv", d, 2 * d),
              (n + "shortcut", "conv", d_in, 2 * d)]
        d_in = 2 * d
    t.append(("decoder_0", "conv", d_in, d_in))
    enc_widths = [2 * cfg.d_out[0]] + [2 * d for d in cfg.d_out[:cfg.num_layers]]  # enc0, samp0..samp4
    feat = d_in
    for j in range(cfg.num_layers):
        skip = enc_widths[-j - 2]
        t.append(("Decoder_layer_%d" % j, "convT", skip + feat, skip))
        feat = skip
    t += [("fc1", "conv", feat, 64), ("fc2", "conv", 64, 32), ("fc", "conv_nobn", 32, cfg.num_classes)]
    return t


def init_params(cfg, num_features: int, seed: int = 0) -> dict:
    """Reference initialisers as numpy fp32 (shared verbatim with the oracle):
    conv kernels truncated_normal(std=sqrt(2/shape[-1])) rounded to 3 decimals (helper_tf_util.py:47-51; shape[-1]
    is Cout for conv2d and Cin for conv2d_transpose), zero biases, glorot-uniform dense kernels, BN gamma=1 beta=0
    moving_mean=0 moving_variance=1."""
    rng = np.random.default_rng(seed)
    p = {}

    def trunc_normal(shape, std):
        x = rng.standard_normal(shape)
        bad = np.abs(x) > 2
        while bad.any():
            x[bad] = rng.standard_normal(int(bad.sum()))
            bad = np.abs(x) > 2
        return (np.round(x * std * 1000) / 1000).astype(np.float32)

    def glorot(shape):
        lim = math.sqrt(6.0 / (shape[0] + shape[1]))
        return rng.uniform(-lim, lim, size=shape).astype(np.float32)

    def bn(scope, c):
        p[scope + "/bn/gamma"] = np.ones(c, np.float32)
        p[scope + "/bn/beta"] = np.zeros(c, np.float32)
        p[scope + "/bn/moving_mean"] = np.zeros(c, np.float32)
        p[scope + "/bn/moving_variance"] = np.ones(c, np.float32)

    for scope, kind, cin, cout in layer_table(cfg, num_features):
        if kind == "dense":
            p[scope + "/kernel"] = glorot((cin, cout))
            p[scope + "/bias"] = np.zeros(cout, np.float32)
            bn(scope, cout)
        elif kind == "att_fc":
            p[scope + "/kernel"] = glorot((cin, cout))
        elif kind == "conv":
            p[scope + "/weights"] = trunc_normal((cin, cout), math.sqrt(2.0 / cout))
            p[scope + "/biases"] = np.zeros(cout, np.float32)
            bn(scope, cout)
        elif kind == "convT":
            p[scope + "/weights"] = trunc_normal((cout, cin), math.sqrt(2.0 / cin))
            p[scope + "/biases"] = np.zeros(cout, np.float32)
            bn(scope, cout)
        elif kind == "conv_nobn":
            p[scope + "/weights"] = trunc_normal((cin, cout), math.sqrt(2.0 / cout))
            p[scope + "/biases"] = np.zeros(cout, np.float32)
    return p


def tf_variable_name(name: str) -> str:
    """Name of variable ``name`` in a checkpoint written by the reference's ``tf.train.Saver(GLOBAL_VARIABLES)``
    (RandLANet.py:56,101-102): everything lives under the ``layers`` variable scope; ``tf.layers.batch_normalization`` is
    unnamed, so inside ``helper_tf_util.conv2d``'s scope it becomes ``<scope>/batch_normalization`` and the one that follows
    ``fc0`` directly under ``layers`` (RandLANet.py:115) is ``layers/batch_normalization``."""
    if name.startswith("fc0/bn/"):
        return "layers/batch_normalization/" + name[len("fc0/bn/"):]
    return "layers/" + name.replace("/bn/", "/batch_normalization/")


def build_pyramid(xyz: torch.Tensor, cfg, side: "torch.cuda.Stream | None" = None, inverse: bool = False,
                  store: "dict | None" = None, locse: bool = False) -> dict:
    """``tf_map`` (runPancreas.py:124-145 / runBraTS.py:140-161) on the device: per level
    ``neigh_idx = knn(xyz, xyz, k_n)``, ``sub = xyz[:, :N//ratio]``, ``sub_idx = neigh_idx[:, :N//ratio]``,
    ``interp_idx = knn(sub, xyz, 1)``.  ``xyz`` is a CUDA ``[B,N,3]`` fp32 tensor; everything stays on the GPU.

    The sub-clouds are prefixes of ``xyz``, so the ten searches do not depend on each other or on the network.  With a
    ``side`` stream only the level-0 neighbour search runs on the current stream; the other nine (and, with ``inverse``,
    the fifteen inverse neighbour lists of the scatter-free backward) are queued on ``side`` and overlap the level-0
    forward, which is HBM-bound while the searches are issue-bound.  The returned dict then carries two callables:
    ``pyramid_ready()`` makes the current stream wait for the remaining indices (Network.inference calls it before the
    first ``random_sample``) and ``inverse_ready()`` for the inverse lists (called before the backward).  All result
    tensors are allocated on the current stream; ``side`` only runs kernels.

    ``locse``: also prepare, per level, what the fused position branch needs from the pyramid alone (``ops.locse_prepare``:
    padded cloud + moments of the LocSE rows) as ``out["locse"]`` -- level 0 on the current stream, the others on ``side``.

    ``store`` (optional): preallocated result tensors -- ``xyz`` (num_layers + 1 clouds; ``xyz[0]`` receives a copy of the
    input), ``neigh_idx``, ``sub_idx``, ``interp_idx`` (num_layers each) and, with ``inverse``, ``inv`` = per level three
    ``(offsets, perm)`` pairs for (neigh_idx, sub_idx, interp_idx).  Everything then runs on the current stream and
    nothing is allocated (the pipelined step of train.py fills such a store one step ahead)."""
    if store is None:
        ops.clear_caches()  # inverse lists cached for a previous pyramid must not be picked up through a recycled buffer
    if store is not None:
        assert side is None
        clouds = store["xyz"]
        clouds[0].copy_(xyz)
        for i in range(cfg.num_layers):
            clouds[i + 1].copy_(clouds[i][:, :clouds[i + 1].shape[1], :])
            if KNN_FUSED_INTERP:
                knn_self_interp_cuda(clouds[i], cfg.k_n, clouds[i + 1].shape[1], store["neigh_idx"][i], store["interp_idx"][i])
            else:
                knn_search_cuda(clouds[i], clouds[i], cfg.k_n, out=store["neigh_idx"][i])
                knn_search_cuda(clouds[i + 1], clouds[i], 1, out=store["interp_idx"][i])
            store["sub_idx"][i].copy_(store["neigh_idx"][i][:, :clouds[i + 1].shape[1], :])
            if "locse" in store:
                ops.locse_prepare(clouds[i], store["neigh_idx"][i], out=store["locse"][i])
            if inverse:
                n, n_sub = clouds[i].shape[1], clouds[i + 1].shape[1]
                for idx, n_src, o in zip((store["neigh_idx"][i], store["sub_idx"][i], store["interp_idx"][i]),
                                         (n, n, n_sub), store["inv"][i]):
                    ops.InverseIndex(idx, n_src, out=o)
        res = dict(xyz=clouds[:-1], neigh_idx=store["neigh_idx"], sub_idx=store["sub_idx"], interp_idx=store["interp_idx"])
        if "locse" in store:
            res["locse"] = store["locse"]
        return res
    out = dict(xyz=[], neigh_idx=[], sub_idx=[], interp_idx=[])
    xyz = xyz.contiguous().float()
    B, dev = xyz.shape[0], xyz.device
    for i in range(cfg.num_layers):
        n = xyz.shape[1]
        n_sub = n // cfg.sub_sampling_ratio[i]
        out["xyz"].append(xyz)
        out["neigh_idx"].append(torch.empty((B, n, cfg.k_n), dtype=torch.int32, device=dev))
        out["sub_idx"].append(torch.empty((B, n_sub, cfg.k_n), dtype=torch.int32, device=dev))
        out["interp_idx"].append(torch.empty((B, n, 1), dtype=torch.int32, device=dev))
        if locse:
            out.setdefault("locse", []).append((torch.empty((B, n, 4), dtype=torch.float32, device=dev),
                                                torch.empty(68, dtype=torch.float32, device=dev)))
        xyz = xyz[:, :n_sub, :].contiguous()
    subs = out["xyz"][1:] + [xyz]

    def search(i):   # neigh_idx (and, fused, interp_idx) of level i
        pts = out["xyz"][i]
        if KNN_FUSED_INTERP:
            knn_self_interp_cuda(pts, cfg.k_n, subs[i].shape[1], out["neigh_idx"][i], out["interp_idx"][i])
        else:
            knn_search_cuda(pts, pts, cfg.k_n, out=out["neigh_idx"][i])

    def level(i):
        pts = out["xyz"][i]
        if i > 0 or side is None:
            search(i)
            if locse:
                ops.locse_prepare(pts, out["neigh_idx"][i], out=out["locse"][i])
        out["sub_idx"][i].copy_(out["neigh_idx"][i][:, :subs[i].shape[1], :])
        if not KNN_FUSED_INTERP:
            knn_search_cuda(subs[i], pts, 1, out=out["interp_idx"][i])

    if side is None:
        for i in range(cfg.num_layers):
            level(i)
        return out
    main = torch.cuda.current_stream(dev)
    search(0)
    if locse:
        ops.locse_prepare(out["xyz"][0], out["neigh_idx"][0], out=out["locse"][0])
    side.wait_stream(main)
    ev = torch.cuda.Event()
    with torch.cuda.stream(side):
        for i in range(cfg.num_layers):
            level(i)
        ev.record(side)
        if inverse:  # the index tensors whose gathers are differentiated: (idx, rows of the gathered tensor per cloud)
            for i in range(cfg.num_layers):
                n, n_sub = out["xyz"][i].shape[1], subs[i].shape[1]
                for idx, n_src in ((out["neigh_idx"][i], n), (out["sub_idx"][i], n), (out["interp_idx"][i], n_sub)):
                    inv = ops.inverse_of(idx, n_src)
                    # allocated while `side` was current but read (and freed) on the main stream
                    inv.offsets.record_stream(main)
                    inv.perm.record_stream(main)
    out["_sub_clouds"] = subs  # the last sub-cloud is read by a search still queued on `side`: it must outlive this call
    out["pyramid_ready"] = lambda: torch.cuda.current_stream(dev).wait_event(ev)
    out["inverse_ready"] = lambda: torch.cuda.current_stream(dev).wait_stream(side)
    return out


class Network(torch.nn.Module):
    """RandLA-Net ``Network`` (RandLANet.py:19-152) on CUDA.  ``config`` mirrors ``helper_tool.Config*``."""

    def __init__(self, config, num_features: int | None = None, seed: int = 0, device="cuda"):
        super().__init__()
        self.config = config
        self.num_features = num_features if num_features is not None else config.num_features
        self._names = {}
        self._stat_names = {}  # reference name -> buffer key of moving_mean / moving_variance (not trained)
        self.vars = torch.nn.ParameterDict()
        self.load_numpy(init_params(config, self.num_features, seed), device)
        self.register_buffer("class_weights", torch.tensor(DP.get_class_weights(config.name).reshape(-1), dtype=torch.float32,
                                                           device=device), persistent=False)

    # -- variables -------------------------------------------------------------------------------
    @staticmethod
    def _key(name: str) -> str:
        return name.replace("/", "__").replace(".", "_")

    @property
    def is_training(self) -> bool:
        """The reference's ``is_training`` placeholder (RandLANet.py:54) follows ``nn.Module.train()`` / ``eval()``."""
        return self.training

    @is_training.setter
    def is_training(self, value: bool):
        self.train(bool(value))

    @property
    def stats(self) -> dict:
        """{reference name: tensor} of the BN moving statistics.  They are registered BUFFERS, so ``state_dict`` /
        ``load_state_dict`` / ``.to()`` cover them like the reference's Saver covers all GLOBAL_VARIABLES (RandLANet.py:101)."""
        return {n: self._buffers[k] for n, k in self._stat_names.items()}

    def load_numpy(self, params: dict, device=None):
        """Inject variables by reference name (the same dict feeds the oracle).  Existing variables are overwritten in
        place (views such as the trainer's flat gradient buffer stay valid)."""
        for name, arr in params.items():
            is_stat = name.endswith("moving_mean") or name.endswith("moving_variance")
            key = "stat__" + self._key(name) if is_stat else self._key(name)
            old = self._buffers.get(key) if is_stat else (self.vars[key] if key in self.vars else None)
            dev = device if device is not None else (old.device if old is not None else "cuda")
            t = torch.as_tensor(np.asarray(arr), dtype=torch.float32).to(dev).contiguous()
            if old is not None and old.shape == t.shape and old.device == t.device:
                with torch.no_grad():
                    old.copy_(t)
            elif is_stat:
                self._stat_names[name] = key
                self.register_buffer(key, t)
            else:
                self._names[name] = key
                self.vars[key] = torch.nn.Parameter(t)

    def load_tf_checkpoint(self, prefix: str, strict: bool = True) -> list:
        """Restore from a checkpoint written by the reference (``snap-<step>``, RandLANet.py:101-102,180-184; restored at
        testPancreas.py:129-132) by VARIABLE NAME: conv kernels ``[1,1,Cin,Cout]`` / ``[1,1,Cout,Cin]`` are squeezed, BN
        names translated (``tf_variable_name``), optimizer slots ignored.  Returns the reference names that were loaded."""
        from .tf_checkpoint import read_checkpoint
        ck = read_checkpoint(prefix)
        params, missing = {}, []
        for name in list(self._names) + list(self._stat_names):
            tf_name = tf_variable_name(name)
            if tf_name not in ck:
                missing.append(tf_name)
                continue
            arr = np.asarray(ck[tf_name])
            want = tuple(self.v(name).shape)
            if arr.ndim == 4 and arr.shape[:2] == (1, 1):
                arr = arr[0, 0]
            if tuple(arr.shape) != want:
                raise ValueError(f"{tf_name}: checkpoint shape {tuple(arr.shape)} != {want}")
            params[name] = arr
        if strict and missing:
            raise KeyError(f"variables missing from {prefix}: {missing[:5]}{' ...' if len(missing) > 5 else ''}")
        self.load_numpy(params)
        return sorted(params)

    def save_tf_checkpoint(self, prefix: str) -> None:
        """Write the variables (and moving statistics) in the reference's checkpoint format and naming."""
        from .tf_checkpoint import write_checkpoint
        out = {}
        for name in list(self._names) + list(self._stat_names):
            a = self.v(name).detach().cpu().numpy()
            if name.endswith("/weights"):
                a = a[None, None]  # helper_tf_util.py:151-152 / :211-212: 1x1 kernels are 4-D
            out[tf_variable_name(name)] = a
        write_checkpoint(prefix, out)

    def v(self, name: str) -> torch.Tensor:
        k = self._stat_names.get(name)
        if k is not None:
            return self._buffers[k]
        return self.vars[self._names[name]]

    def named_variables(self):
        """(reference name, tensor) for every trainable variable."""
        return [(n, self.vars[k]) for n, k in self._names.items()]

    def grads_numpy(self) -> dict:
        return {n: (t.grad.detach().cpu().numpy() if t.grad is not None else None) for n, t in self.named_variables()}

    # -- layers (helper_tf_util.py) --------------------------------------------------------------
    def _bn_stats(self, scope, mean, var, count, is_training, fused_4d=True):
        """(mean, var, moving): batch statistics in training plus the moving-average buffers to update (momentum 0.99,
        run with the step like UPDATE_OPS at RandLANet.py:90,163); moving statistics at inference."""
        mm, mv = self.v(scope + "/bn/moving_mean"), self.v(scope + "/bn/moving_variance")
        if not is_training:
            return mm, mv, None
        # TF's fused kernel (4-D NHWC inputs) feeds the UNBIASED variance to the moving average
        unbias = count / max(count - 1, 1) if fused_4d else 1.0
        return mean, var, ((mm, mv, unbias) if torch.is_grad_enabled() else None)

    def conv2d(self, x, scope, bn=True, is_training=True, activation=True, transpose=False, dense_names=False):
        """1x1 conv (+bias) [-> BN(0.99, 1e-6)] [-> LeakyReLU(0.2)]  (helper_tf_util.py:115-170 / :173-250)."""
        wname, bname = ("/kernel", "/bias") if dense_names else ("/weights", "/biases")
        w = self.v(scope + wname)
        if transpose:
            w = w.t()
        b = self.v(scope + bname)
        if not bn:
            if activation:
                raise NotImplementedError("activation without batch norm is not used by PointSegment")
            return ops.linear(x, w, b)
        rows_n = x.numel() // x.shape[-1]
        if is_training:
            y, mean, var = ops.linear(x, w, b, want_stats=True, zero_bias_grad=True, defer_stats=True)
        else:
            y = ops.linear(x, w, b)
            mean = var = None
        mean, var, moving = self._bn_stats(scope, mean, var, rows_n, is_training, fused_4d=not dense_names)
        return ops.bn_act(y, mean, var, self.v(scope + "/bn/gamma"), self.v(scope + "/bn/beta"),
                          slope=ops.LEAKY_SLOPE if activation else 1.0, training=is_training, moving=moving)

    # -- LFA ops (RandLANet.py:337-401) ----------------------------------------------------------
    @staticmethod
    def gather_neighbour(pc, neighbor_idx):
        return ops.gather_neighbour(pc, neighbor_idx)

    @staticmethod
    def relative_pos_encoding(xyz, neigh_idx):
        return ops.relative_pos_encoding(xyz, neigh_idx)

    @staticmethod
    def random_sample(feature, pool_idx):
        return ops.random_sample(feature, pool_idx)

    @staticmethod
    def nearest_interpolation(feature, interp_idx):
        return ops.nearest_interpolation(feature, interp_idx)

    def att_pooling(self, feature_set, d_out, name, is_training):
        f_agg = ops.att_pool(feature_set, self.v(name + "fc/kernel"))
        return self.conv2d(f_agg, name + "mlp", True, is_training, True)

    def _linear_stats(self, x, scope, is_training):
        """1x1 conv + batch statistics, without the BN/activation (fused into the consumer)."""
        w, b = self.v(scope + "/weights"), self.v(scope + "/biases")
        rows_n = x.numel() // x.shape[-1]
        if is_training:
            y, mean, var = ops.linear(x, w, b, want_stats=True, zero_bias_grad=True, defer_stats=True)
        else:
            y, mean, var = ops.linear(x, w, b), None, None
        mean, var, moving = self._bn_stats(scope, mean, var, rows_n, is_training)
        return y, mean, var, self.v(scope + "/bn/gamma"), self.v(scope + "/bn/beta"), moving

    def building_block(self, xyz, feature, neigh_idx, d_out, name, is_training, locse_pre=None):
        """RandLANet.py:323-335.  Same dataflow.  The position branch (LocSE -> mlp1 -> BN -> LeakyReLU) runs as recompute
        kernels (csrc/locse_mlp.cu); the two tf.concat's are produced in place by ops.lfa_concat (gather into the left half,
        BN + LeakyReLU of the position MLP into the right half) -- or, at the 16-channel level, not at all: the att16 kernels
        take the two halves as separate tensors (32-byte half rows inside 64-byte rows waste half of every DRAM burst)."""
        scope = name + "mlp1"
        h = self.v(scope + "/weights").shape[1]
        K = neigh_idx.shape[-1]
        fused = ops.locse_mlp_supported(K, h)
        split = fused and ops.att_pool_split_supported(K, 2 * h)
        if fused:   # LocSE + mlp1 + BN + LeakyReLU recomputed inside one kernel each way
            rows_n = neigh_idx.numel()
            f_concat, f_xyz = ops.locse_mlp_concat(
                xyz, None if split else feature.squeeze(2), neigh_idx, self.v(scope + "/weights"), self.v(scope + "/biases"),
                self.v(scope + "/bn/gamma"), self.v(scope + "/bn/beta"), is_training, self.v(scope + "/bn/moving_mean"),
                self.v(scope + "/bn/moving_variance"), rows_n / max(rows_n - 1, 1), is_training and torch.is_grad_enabled(),
                pre=locse_pre)
        else:
            f_xyz = self.relative_pos_encoding(xyz, neigh_idx)
            y, m, v, g, b, mv = self._linear_stats(f_xyz, scope, is_training)
            f_concat, f_xyz = ops.lfa_concat(feature.squeeze(2), neigh_idx, y, m, v, g, b, is_training, mv, need_fxyz=True)
        if split:   # f_concat is the first alias of f_xyz here
            f_agg = ops.att_pool_split(self.gather_neighbour(feature.squeeze(2), neigh_idx), f_concat,
                                       self.v(name + "att_pooling_1fc/kernel"))
            f_pc_agg = self.conv2d(f_agg, name + "att_pooling_1mlp", True, is_training, True)
            f_xyz = self.conv2d(f_xyz, name + "mlp2", True, is_training, True)
            f_agg = ops.att_pool_split(self.gather_neighbour(f_pc_agg.squeeze(2), neigh_idx), f_xyz,
                                       self.v(name + "att_pooling_2fc/kernel"))
            return self.conv2d(f_agg, name + "att_pooling_2mlp", True, is_training, True)
        f_pc_agg = self.att_pooling(f_concat, d_out // 2, name + "att_pooling_1", is_training)
        y, m, v, g, b, mv = self._linear_stats(f_xyz, name + "mlp2", is_training)
        f_concat, _ = ops.lfa_concat(f_pc_agg.squeeze(2), neigh_idx, y, m, v, g, b, is_training, mv, need_fxyz=False)
        return self.att_pooling(f_concat, d_out, name + "att_pooling_2", is_training)

    def dilated_res_block(self, feature, xyz, neigh_idx, d_out, name, is_training, locse_pre=None):
        f_pc = self.conv2d(feature, name + "mlp1", True, is_training)
        f_pc = self.building_block(xyz, f_pc, neigh_idx, d_out, name + "LFA", is_training, locse_pre)
        # mlp2 / shortcut: BN without activation, then leaky_relu(sum)  (RandLANet.py:317-321), one fused kernel
        outs = []
        for x, scope in ((f_pc, name + "mlp2"), (feature, name + "shortcut")):
            w, b = self.v(scope + "/weights"), self.v(scope + "/biases")
            rows_n = x.numel() // x.shape[-1]
            if is_training:
                y, mean, var = ops.linear(x, w, b, want_stats=True, zero_bias_grad=True, defer_stats=True)
            else:
                y, mean, var = ops.linear(x, w, b), None, None
            mean, var, moving = self._bn_stats(scope, mean, var, rows_n, is_training)
            outs.append((y, mean, var, self.v(scope + "/bn/gamma"), self.v(scope + "/bn/beta"), moving))
        (y1, m1, v1, g1, b1, mv1), (y2, m2, v2, g2, b2, mv2) = outs
        return ops.bn_act(y1, m1, v1, g1, b1, slope=ops.LEAKY_SLOPE, training=is_training, moving=mv1,
                          y2=y2, mean2=m2, var2=v2, gamma2=g2, beta2=b2, moving2=mv2)

    def inference(self, inputs, is_training, dropout_mask=None):
        """RandLANet.py:110-152.  ``inputs``: dict(xyz, neigh_idx, sub_idx, interp_idx: lists of 5; features [B,N,F])."""
        cfg = self.config
        feature = self.conv2d(inputs["features"], "fc0", True, is_training, True, dense_names=True)
        feature = feature.unsqueeze(2)
        f_encoder_list = []
        pre = inputs.get("locse") if is_training else None   # (padded cloud, LocSE moments) per level, from build_pyramid
        for i in range(cfg.num_layers):
            f_encoder_i = self.dilated_res_block(feature, inputs["xyz"][i], inputs["neigh_idx"][i], cfg.d_out[i],
                                                 "Encoder_layer_" + str(i), is_training, pre[i] if pre is not None else None)
            if i == 0 and "pyramid_ready" in inputs:
                inputs["pyramid_ready"]()  # the rest of the index pyramid was built on a side stream (build_pyramid)
            f_sampled_i = self.random_sample(f_encoder_i, inputs["sub_idx"][i])
            feature = f_sampled_i
            if i == 0:
                f_encoder_list.append(f_encoder_i)
            f_encoder_list.append(f_sampled_i)
        feature = self.conv2d(f_encoder_list[-1], "decoder_0", True, is_training)
        f_decoder_list = []
        for j in range(cfg.num_layers):
            f_interp_i = self.nearest_interpolation(feature, inputs["interp_idx"][-j - 1])
            f_decoder_i = self.conv2d(torch.cat([f_encoder_list[-j - 2], f_interp_i], dim=3), "Decoder_layer_" + str(j),
                                      True, is_training, transpose=True)
            feature = f_decoder_i
            f_decoder_list.append(f_decoder_i)
        f_layer_fc1 = self.conv2d(f_decoder_list[-1], "fc1", True, is_training)
        f_layer_fc2 = self.conv2d(f_layer_fc1, "fc2", True, is_training)
        if is_training:
            if dropout_mask is None:  # tf.nn.dropout(keep_prob=0.5) (helper_tf_util.py:571-573): inverted scaling, x / keep
                f_layer_drop = torch.nn.functional.dropout(f_layer_fc2, p=0.5, training=True)  # one fused kernel each way
            else:                      # injected keep-mask (parity tests share it with the oracle)
                f_layer_drop = f_layer_fc2 * (dropout_mask.to(f_layer_fc2.dtype) * 2.0)
        else:
            f_layer_drop = f_layer_fc2
        f_layer_fc3 = self.conv2d(f_layer_drop, "fc", False, is_training, activation=False)
        return f_layer_fc3.squeeze(2)

    def forward(self, inputs, dropout_mask=None):
        return self.inference(inputs, self.training, dropout_mask)

    def get_loss(self, logits, labels):
        """RandLANet.py:62-84,267-274 (no ignored labels for Pancreas/BraTS): class-weighted CE, mean over points."""
        C = logits.shape[-1]
        logits = logits.reshape(-1, C)
        labels = labels.reshape(-1).long()
        w = self.class_weights[labels]
        return (torch.nn.functional.cross_entropy(logits, labels, reduction="none") * w).mean()
