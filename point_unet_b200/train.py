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
GRAPH_PRIORITY = int(_os.environ.get("PU_GRAPH_PRIORITY", "1")) != 0


class _Slot:
    """Static storage of one batch, its index pyramid and the inverse neighbour lists of the backward -- all of them
    views of ONE flat byte buffer, so handing a finished slot over to the training step is a single device copy."""

    def __init__(self, cfg, B, N, n_feat, labels_dtype, device):
        spec, n = [], N
        i32, f32 = torch.int32, torch.float32
        for i in range(cfg.num_layers):
            n_sub = n // cfg.sub_sampling_ratio[i]
            spec += [(f"xyz{i}", (B, n, 3), f32), (f"neigh{i}", (B, n, cfg.k_n), i32), (f"sub{i}", (B, n_sub, cfg.k_n), i32),
                     (f"interp{i}", (B, n, 1), i32), (f"xyz4_{i}", (B, n, 4), f32), (f"locse_mom{i}", (68,), f32),
                     (f"inv_neigh_off{i}", (B * n + 1,), i32), (f"inv_neigh_perm{i}", (B * n * cfg.k_n,), i32),
                     (f"inv_sub_off{i}", (B * n + 1,), i32), (f"inv_sub_perm{i}", (B * n_sub * cfg.k_n,), i32),
                     (f"inv_interp_off{i}", (B * n_sub + 1,), i32), (f"inv_interp_perm{i}", (B * n,), i32)]
            n = n_sub
        spec += [(f"xyz{cfg.num_layers}", (B, n, 3), f32), ("features", (B, N, n_feat), f32), ("labels", (B, N), labels_dtype)]
        offs, total = [], 0
        for _, shape, dt in spec:
            offs.append(total)
            nbytes = int(np.prod(shape)) * torch.empty((), dtype=dt).element_size()
            total += (nbytes + 255) // 256 * 256
        self.flat = torch.zeros(total, dtype=torch.uint8, device=device)
        self.t = {}
        for (name, shape, dt), o in zip(spec, offs):
            nbytes = int(np.prod(shape)) * torch.empty((), dtype=dt).element_size()
            self.t[name] = self.flat[o:o + nbytes].view(dt).view(shape)
        L = range(cfg.num_layers)
        self.store = dict(xyz=[self.t[f"xyz{i}"] for i in range(cfg.num_layers + 1)],
                          neigh_idx=[self.t[f"neigh{i}"] for i in L], sub_idx=[self.t[f"sub{i}"] for i in L],
                          interp_idx=[self.t[f"interp{i}"] for i in L],
                          locse=[(self.t[f"xyz4_{i}"], self.t[f"locse_mom{i}"]) for i in L],
                          inv=[[(self.t[f"inv_{k}_off{i}"], self.t[f"inv_{k}_perm{i}"]) for k in ("neigh", "sub", "interp")]
                               for i in L])
        self.features, self.labels = self.t["features"], self.t["labels"]

    def pyramid(self):
        st = self.store
        return dict(xyz=st["xyz"][:-1], neigh_idx=st["neigh_idx"], sub_idx=st["sub_idx"], interp_idx=st["interp_idx"],
                    locse=st["locse"])

    def register_inverse(self):
        """Point ops.inverse_of at this slot's lists (the cache is cleared at the end of every step)."""
        st = self.store
        for i, inv in enumerate(st["inv"]):
            n, n_sub = st["xyz"][i].shape[1], st["xyz"][i + 1].shape[1]
            for idx, n_src, (off, perm) in zip((st["neigh_idx"][i], st["sub_idx"][i], st["interp_idx"][i]), (n, n, n_sub), inv):
                ops.register_inverse(idx, n_src, off, perm)


class Trainer:
    def __init__(self, config, num_features=None, seed=0, device=None, lr=None, world_size=1, storage=None):
        """``storage``: "fp32" (default: the 1e-3 parity path) or "bf16" -- store the pre-normalisation activations as
        bfloat16 (ops.set_storage; stated tolerance 2e-2, tests/test_bf16_storage_gpu.py).  Process-wide switch."""
        if storage is not None:
            ops.set_storage(storage)
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
        # The learning rate is a DEVICE scalar: a Python float would be baked into the captured CUDA graph, a tensor is
        # read by every replay, so the reference's per-epoch decay (RandLANet.py:190-193) works on the graph path too.
        self.lr = torch.tensor(float(lr if lr is not None else config.learning_rate), dtype=torch.float32, device=self.device)
        self.opt = torch.optim.Adam(self.params, lr=self.lr, betas=(0.9, 0.999), eps=1e-8, fused=True, capturable=True)
        self.world_size = world_size
        self._pinned = {}
        self._dev = {}
        self._h2d_done = {}
        # [loss, tcgen05 error flag] of the last step, written on the device at the end of every step and fetched with ONE
        # small D2H copy by the host-facing entry points
        self._report_dev = torch.zeros(2, dtype=torch.float32, device=self.device)
        self._report_host = torch.zeros(2, dtype=torch.float32, pin_memory=True)
        self._report_ev = torch.cuda.Event()
        self._copy_stream = None
        self._stage_slots = None
        self._stage_ev = [None, None]
        self._stage_k = 0
        self._graph = None
        self._side = None
        self._side2 = None
        self.pipelined = False
        self.dropout_mask = None  # fixed [B,N,1,32] bool mask instead of a fresh tf.nn.dropout draw per step (tests)

    # -- host staging ----------------------------------------------------------------------------
    def _stage(self, name, arr):
        """numpy / CPU tensor -> (pinned buffer ->) device, async on the current stream.  A tensor that already lives in
        pinned memory is copied straight from where it is (no pinned -> pinned memcpy); pageable input goes through a
        staging buffer that is rewritten only after the previous H2D copy out of it has completed."""
        t = torch.from_numpy(np.ascontiguousarray(arr)) if isinstance(arr, np.ndarray) else arr
        if t.is_cuda:
            return t
        dev = self._dev.get(name)
        if dev is None or dev.shape != t.shape or dev.dtype != t.dtype:
            dev = self._dev[name] = torch.empty(t.shape, dtype=t.dtype, device=self.device)
        src = t
        if not t.is_pinned():
            pin = self._pinned.get(name)
            if pin is None or pin.shape != t.shape or pin.dtype != t.dtype:
                pin = self._pinned[name] = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
            ev = self._h2d_done.get(name)
            if ev is not None:
                ev.synchronize()       # the previous copy out of the staging buffer may still be queued
            pin.copy_(t)
            src = pin
        dev.copy_(src, non_blocking=True)
        ev = self._h2d_done.get(name)
        if ev is None:
            ev = self._h2d_done[name] = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        return dev

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
        # tf_map on the GPU; levels 1-4 and the inverse lists of the backward are built on a side stream under the
        # level-0 forward (PU_OVERLAP=0: everything on one stream)
        pyr = build_pyramid(xyz, self.cfg, side=self._side_stream() if OVERLAP else None, inverse=True, locse=True)
        return self._train_on(pyr, xyz, features, labels, dropout_mask)

    def _train_on(self, pyr, xyz, features, labels, dropout_mask=None):
        """Forward, loss, backward, gradient all-reduce and Adam on a finished index pyramid."""
        net = self.net
        if dropout_mask is None:
            dropout_mask = self.dropout_mask
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
            ops._wgrad_keep.clear()
        if self.world_size > 1:
            self.bucket.all_reduce_mean()                          # gradients only, NCCL over NVLink
        self.opt.step()
        ops.clear_caches()
        loss = loss.detach()
        self._report_dev[0].copy_(loss)
        self._report_dev[1].copy_(ops.tc_error_flag(self.device)[0])
        return loss

    # -- learning-rate schedule, error reporting ----------------------------------------------------------
    def decay_lr(self, factor=None):
        """``lr *= lr_decays[epoch]`` after every epoch (RandLANet.py:190-193; 0.95 for both configs).  In place on the device
        scalar the optimiser reads, so eager steps AND graph replays pick it up."""
        self.lr.mul_(float(factor if factor is not None else self.cfg.lr_decays))
        return self

    end_epoch = decay_lr

    def _fetch_report(self):
        """(loss, error flag) of the last step: one 8-byte D2H copy into pinned memory, then wait for it."""
        self._report_host.copy_(self._report_dev, non_blocking=True)
        self._report_ev.record(torch.cuda.current_stream(self.device))
        self._report_ev.synchronize()
        return float(self._report_host[0]), float(self._report_host[1])

    def check_errors(self, flag_value=None):
        """Raise if a tcgen05 kernel's bounded mbarrier wait timed out since the last check (the kernels then carried on
        with whatever was in tensor / shared memory, so the step's numbers are garbage); the flag is reset."""
        if flag_value is None:
            flag_value = float(ops.tc_error_flag(self.device).item())
        if flag_value != 0.0:
            ops.tc_error_flag(self.device).zero_()
            self._report_dev[1].zero_()
            from ._lib import PointUnetError
            raise PointUnetError("a tcgen05 pipeline barrier timed out during the last step: its results are invalid")

    # -- CUDA graph: ~1300 kernel launches per step recorded once, replayed with one host call -------
    def capture_step(self, xyz, features, labels, warmup=3, pipelined=False):
        """Record one optimisation step for inputs of this shape into a CUDA graph.  ``xyz/features/labels`` are device
        tensors used for the warm-up (workspaces, kernel attributes and the allocator pool are set up outside the
        capture); afterwards ``train_step_graph`` copies a new batch into the static input buffers and replays.
        The warm-up steps ARE optimisation steps (they update the weights).

        ``pipelined``: software-pipeline the input side one step ahead, the way the reference hides ``tf_map`` behind
        the training step with ``tf.data ... prefetch`` (runPancreas.py:124-171).  One replay then does two things side by
        side: the main stream trains on batch i, whose index pyramid and inverse lists were finished by the previous
        replay, while a side stream builds the pyramid of batch i+1 -- the batch handed to this call -- into the other
        slot.  Per replay the work is the same (one pyramid, one forward/backward/Adam); the ten searches just leave the
        critical path.  ``train_step_graph`` therefore returns the loss of the batch submitted ONE CALL EARLIER (the first
        replay trains on the capture batch)."""
        from . import _lib
        self._gx, self._gf, self._gl = xyz.clone(), features.clone(), labels.clone()
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(max(int(warmup), 1)):
                self.train_step_device(self._gx, self._gf, self._gl)
            if pipelined:
                B, N = xyz.shape[0], xyz.shape[1]
                self._slots = [_Slot(self.cfg, B, N, features.shape[-1], labels.dtype, self.device) for _ in range(2)]
                self._fill_slot(self._slots[1], self._gx, self._gf, self._gl)  # prologue: the first replay trains on this
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        graph = torch.cuda.CUDAGraph()
        n0 = _lib.launch_count()
        # The forward/backward chain is captured from a HIGH-priority stream, the side streams (pyramid, inverse lists,
        # weight gradients) have the default, lowest priority: kernel nodes inherit it, so whenever an SM frees up the
        # block scheduler places the chain's CTAs first and the side work fills what is left.
        hi = torch.cuda.Stream(device=self.device, priority=-5) if GRAPH_PRIORITY else None
        with torch.cuda.graph(graph, stream=hi):
            if pipelined:
                self._gloss = self._pipelined_step()
            else:
                self._gloss = self.train_step_device(self._gx, self._gf, self._gl)
        self.graph_launches = int(_lib.launch_count() - n0)  # our kernels inside one replay (torch's own come on top)
        self._graph = graph
        self.pipelined = bool(pipelined)
        return self

    def _fill_slot(self, slot, xyz, features, labels):
        """Batch -> slot on the current stream: copies, the ten searches of tf_map, the fifteen inverse lists."""
        slot.features.copy_(features)
        slot.labels.copy_(labels)
        build_pyramid(xyz, self.cfg, inverse=True, store=slot.store)

    def _pipelined_step(self):
        cur, nxt = self._slots
        main = torch.cuda.current_stream(self.device)
        cur.flat.copy_(nxt.flat)                 # batch i with its pyramid and inverse lists: one 200 MB device copy
        side = self._side_stream()
        side.wait_stream(main)
        with torch.cuda.stream(side):            # batch i+1 (just written to the static input buffers)
            self._fill_slot(nxt, self._gx, self._gf, self._gl)
        cur.register_inverse()
        loss = self._train_on(cur.pyramid(), cur.store["xyz"][0], cur.features, cur.labels)
        main.wait_stream(side)
        return loss

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
        (device -> host read) and raises if a tensor-core pipeline reported a barrier timeout.

        Without a captured graph: stage, train, read back, in that order.  With a graph the host side is overlapped the way
        the reference overlaps its input pipeline (``tf.data ... prefetch``, runPancreas.py:158-166): call k queues the
        host -> device copy of batch k on a COPY stream into staging slot k % 2 and, on the main stream, replays the graph
        on the batch that finished copying during call k-1 -- so the copy of batch k runs under the replay.  The value
        returned is therefore the loss of an EARLIER batch: one call back for a plain graph, two for a pipelined one
        (``capture_step(pipelined=True)`` already trains one batch behind the one it indexes); the first calls train on the
        capture batch.  Pinned input tensors are read asynchronously: leave them unchanged until the next call returns."""
        if self._graph is None or tuple(xyz.shape) != tuple(self._gx.shape) or \
                tuple(features.shape) != tuple(self._gf.shape):
            x = self._stage("xyz", xyz)
            f = self._stage("features", features)
            l = self._stage("labels", labels)
            self.train_step_device(x, f, l)
            loss, flag = self._fetch_report()
            self.check_errors(flag)
            return loss
        main = torch.cuda.current_stream(self.device)
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
            self._stage_slots = [(self._gx.clone(), self._gf.clone(), self._gl.clone()) for _ in range(2)]
        k = self._stage_k
        cur, prev = self._stage_slots[k % 2], self._stage_slots[(k + 1) % 2]
        cs = self._copy_stream
        cs.wait_stream(main)   # slot k % 2 was read by the device copy of call k - 1, queued on `main`
        with torch.cuda.stream(cs):
            for dst, src, name in zip(cur, (xyz, features, labels), ("xyz", "features", "labels")):
                t = torch.from_numpy(np.ascontiguousarray(src)) if isinstance(src, np.ndarray) else src
                if not t.is_pinned():   # pageable input: through a per-slot pinned buffer (rewritten only once its copy is done)
                    key = (name, k % 2)
                    pin = self._pinned.get(key)
                    if pin is None or pin.shape != t.shape or pin.dtype != t.dtype:
                        pin = self._pinned[key] = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
                    if self._stage_ev[k % 2] is not None:
                        self._stage_ev[k % 2].synchronize()
                    pin.copy_(t)
                    t = pin
                dst.copy_(t, non_blocking=True)
            ev = self._stage_ev[k % 2] or torch.cuda.Event()
            ev.record(cs)
            self._stage_ev[k % 2] = ev
        if self._stage_ev[(k + 1) % 2] is not None:
            main.wait_event(self._stage_ev[(k + 1) % 2])       # batch k - 1 has landed (first call: the capture batch)
        self.train_step_graph(*prev)
        self._stage_k = k + 1
        loss, flag = self._fetch_report()
        self.check_errors(flag)
        return loss

    def flush_inputs(self):
        """Wait until every queued host -> device copy has completed (pinned inputs may be modified again)."""
        for ev in list(self._h2d_done.values()) + [e for e in self._stage_ev if e is not None]:
            ev.synchronize()

    @torch.no_grad()
    def predict(self, xyz, features):
        """Test-mode forward (moving BN statistics, no dropout) -> softmax probabilities [B,N,C] on the device
        (testPancreas.py:133-134 ``prob_logits``)."""
        x = self._stage("xyz", xyz)
        f = self._stage("features", features)
        pyr = build_pyramid(x, self.cfg, side=self._side_stream() if OVERLAP else None)
        logits = self.net.inference(dict(pyr, features=torch.cat([x, f], dim=-1)), False)
        ops.clear_caches()
        probs = torch.softmax(logits, dim=-1)
        self.check_errors()   # synchronises: garbage from a timed-out tensor-core pipeline must not leave this call
        return probs

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

    @torch.no_grad()
    def predict_to_labels(self, xyz, features, xyz_origin, volume_shape, point_idx=None, remap=None):
        """Test-mode fusion all the way to the segmentation the reference writes as NIfTI
        (utils/genSegmentationPancreas.py:67-77 / genSegmentationBraTS.py:67-78): uint8 label volumes ``[Z,Y,X]`` =
        argmax over classes of the scattered probabilities, produced on the device without the dense probability volume.
        ``remap=(3, 4)`` for BraTS."""
        probs = self.predict(xyz, features)
        out = []
        for b in range(probs.shape[0]):
            xo = xyz_origin[b]
            xo = torch.as_tensor(np.ascontiguousarray(xo)).to(self.device) if not torch.is_tensor(xo) else xo.to(self.device)
            pi = None
            if point_idx is not None:
                pi = point_idx[b]
                pi = torch.as_tensor(np.ascontiguousarray(pi)).to(self.device) if not torch.is_tensor(pi) else pi.to(self.device)
            out.append(ops.point2label(probs[b], xo, volume_shape, pi, remap=remap))
        return out
