"""Training / inference steps of PointSegment on B200 -- the public API a user of the reference switches to.

Reference flow replaced (PointSegment/RandLANet.py:156-206 ``Network.train`` + runPancreas.py:124-171 input
pipeline): per step the reference runs ``tf_map`` (10 nanoflann KNN calls on the CPU), copies 24 tensors to the
GPU, runs forward + TF autodiff + Adam, and fetches logits.  Here one process per GPU does all of it on the
device: host batch -> pinned staging -> H2D -> GPU index pyramid (csrc/knn.cu) -> forward/backward on the
hand-written kernels -> (N>1: NCCL all-reduce of the flat gradient buffer, gradients only) -> Adam.

Data parallelism: batch-sharded, batch-norm statistics stay per replica, the loss is a
local mean, so averaged gradients equal the global-batch mean gradient.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from . import ops
from .parallel import FlatGradBucket
from .RandLANet import Network, build_pyramid

import os as _os

OVERLAP = int(_os.environ.get("PU_OVERLAP", "1")) != 0


class Trainer:
    def __init__(self, config, num_features=None, seed=0, device=None, lr=None, world_size=1):
        self.device = torch.device(device if device is not None else "cuda")
        self.cfg = config
        self.net = Network(config, num_features, seed=seed, device=self.device)
        self.params = [t for _, t in self.net.named_variables()]
        # one flat gradient buffer: every .grad is a view of it (a single all-reduce, no bucketing copies)
        self.bucket = FlatGradBucket(self.params, self.device)
        self.flat_grad = self.bucket.flat
        # tf.train.AdamOptimizer defaults (RandLANet.py:88): beta (0.9, 0.999), eps 1e-8
        # capturable: the step counter lives on the device, so the whole step (pyramid, forward, backward, all-reduce,
        # Adam) can be recorded once into a CUDA graph and replayed (capture_step / train_step_graph below)
        self.opt = torch.optim.Adam(self.params, lr=lr if lr is not None else config.learning_rate, betas=(0.9, 0.999),
                                    eps=1e-8, fused=True, capturable=True)
        self.world_size = world_size
        self._pinned = {}
        self._dev = {}
        self._graph = None
        self._side = None
        self._side2 = None

    # -- host staging ----------------------------------------------------------------------------
    def _stage(self, name, arr):
        """numpy / CPU tensor -> pinned buffer -> device (async on the current stream)."""
        t = torch.from_numpy(np.ascontiguousarray(arr)) if isinstance(arr, np.ndarray) else arr
        if t.is_cuda:
            return t
        pin = self._pinned.get(name)
        if pin is None or pin.shape != t.shape or pin.dtype != t.dtype:
            pin = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
            self._pinned[name] = pin
            self._dev[name] = torch.empty(t.shape, dtype=t.dtype, device=self.device)
        if t.data_ptr() != pin.data_ptr():
            pin.copy_(t)
        self._dev[name].copy_(pin, non_blocking=True)
        return self._dev[name]

    def pin_batch(self, xyz, features, labels):
        """Pre-place a host batch in pinned memory (what a data-loader worker would hand over)."""
        out = {}
        for name, arr in (("xyz", xyz), ("features", features), ("labels", labels)):
            t = torch.from_numpy(np.ascontiguousarray(arr))
            out[name] = torch.empty(t.shape, dtype=t.dtype, pin_memory=True).copy_(t)
        return out

    def _wgrad_stream(self):
        if self._side2 is None:
            self._side2 = torch.cuda.Stream(device=self.device)
        return self._side2

    def _side_stream(self):
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.device)
        return self._side

    # -- steps -----------------------------------------------------------------------------------
    def train_step_device(self, xyz, features, labels, dropout_mask=None):
        """One optimisation step on device-resident inputs: xyz [B,N,3] f32, features [B,N,F-3] f32, labels [B,N]."""
        net = self.net
        # tf_map on the GPU; levels 1-4 and the inverse lists of the backward are built on a side stream under the
        # level-0 forward (PU_OVERLAP=0: everything on one stream)
        pyr = build_pyramid(xyz, self.cfg, side=self._side_stream() if OVERLAP else None, inverse=True)
        inputs = dict(pyr, features=torch.cat([xyz, features], dim=-1))  # runPancreas.py:125
        self.flat_grad.zero_()
        logits = net.inference(inputs, True, dropout_mask)
        loss = net.get_loss(logits, labels)
        if "inverse_ready" in pyr:
            pyr["inverse_ready"]()
        ops.GRAD_SINK = True   # kernels add parameter gradients straight into the flat buffer's views (ops._sink)
        ops.WGRAD_STREAM = self._wgrad_stream() if OVERLAP else None  # weight gradients next to the dgrad chain
        try:
            loss.backward()
            ops.wgrad_join(self.device)
        finally:
            ops.GRAD_SINK = False
            ops.WGRAD_STREAM = None
        if self.world_size > 1:
            self.bucket.all_reduce_mean()                          # gradients only, NCCL over NVLink
        self.opt.step()
        ops.clear_caches()
        return loss.detach()

    # -- CUDA graph: ~1300 kernel launches per step recorded once, replayed with one host call -------
    def capture_step(self, xyz, features, labels, warmup=3):
        """Record one optimisation step for inputs of this shape into a CUDA graph.  ``xyz/features/labels`` are device
        tensors used for the warm-up (workspaces, kernel attributes and the allocator pool are set up outside the
        capture); afterwards ``train_step_graph`` copies a new batch into the static input buffers and replays.
        The warm-up steps ARE optimisation steps (they update the weights)."""
        from . import _lib
        self._gx, self._gf, self._gl = xyz.clone(), features.clone(), labels.clone()
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(max(int(warmup), 1)):
                self.train_step_device(self._gx, self._gf, self._gl)
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        graph = torch.cuda.CUDAGraph()
        n0 = _lib.launch_count()
        with torch.cuda.graph(graph):
            self._gloss = self.train_step_device(self._gx, self._gf, self._gl)
        self.graph_launches = int(_lib.launch_count() - n0)  # our kernels inside one replay (torch's own come on top)
        self._graph = graph
        return self

    def train_step_graph(self, xyz, features, labels):
        """Replay the captured step on a new device-resident batch of the captured shape; returns the loss tensor."""
        if self._graph is None:
            raise RuntimeError("capture_step() first")
        self._gx.copy_(xyz, non_blocking=True)
        self._gf.copy_(features, non_blocking=True)
        self._gl.copy_(labels, non_blocking=True)
        self._graph.replay()
        return self._gloss

    def train_step(self, xyz, features, labels):
        """Public end-to-end step on HOST buffers (numpy or pinned CPU tensors); returns the loss as a float
        (device -> host read).  Uses the captured graph when there is one for this shape."""
        x = self._stage("xyz", xyz)
        f = self._stage("features", features)
        l = self._stage("labels", labels)
        if self._graph is not None and x.shape == self._gx.shape and f.shape == self._gf.shape:
            return float(self.train_step_graph(x, f, l).item())
        return float(self.train_step_device(x, f, l).item())

    @torch.no_grad()
    def predict(self, xyz, features):
        """Test-mode forward (moving BN statistics, no dropout) -> softmax probabilities [B,N,C] on the device
        (testPancreas.py:133-134 ``prob_logits``)."""
        x = self._stage("xyz", xyz)
        f = self._stage("features", features)
        pyr = build_pyramid(x, self.cfg, side=self._side_stream() if OVERLAP else None)
        logits = self.net.inference(dict(pyr, features=torch.cat([x, f], dim=-1)), False)
        ops.clear_caches()
        return torch.softmax(logits, dim=-1)

    @torch.no_grad()
    def predict_to_volume(self, xyz, features, xyz_origin, volume_shape, point_idx=None):
        """Test-mode fusion (testPancreas.py:141-202 / testBraTS.py:155-232) for ONE cloud per batch row: softmax
        probabilities scattered through the saved integer voxel coordinates into dense ``[Z,Y,X,C]`` volumes.
        ``xyz_origin`` is ``[B, n, 3]`` (or a list), ``volume_shape = (Z, X, Y, C)`` as the reference allocates it."""
        probs = self.predict(xyz, features)
        vols = []
        for b in range(probs.shape[0]):
            xo = xyz_origin[b]
            xo = torch.as_tensor(np.ascontiguousarray(xo)).to(self.device) if not torch.is_tensor(xo) else xo.to(self.device)
            pi = None
            if point_idx is not None:
                pi = point_idx[b]
                pi = torch.as_tensor(np.ascontiguousarray(pi)).to(self.device) if not torch.is_tensor(pi) else pi.to(self.device)
            vols.append(ops.point2prod(probs[b], xo, volume_shape, pi))
        return vols
