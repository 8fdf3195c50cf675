"""Host-side mirror of ``PointSegment/helper_tool.py`` for the hot path: configs + ``DataProcessing.knn_search``.

``DataProcessing.knn_search(support_pts, query_pts, k)`` keeps the reference signature and result
(``helper_tool.py:84-94``: int32 ``[B, N2, k]``, neighbours ascending by distance) but runs the sm_100a
kernel in ``csrc/knn.cu`` through the C-ABI ``pu_knn_batch``.  numpy in -> numpy out (host buffers are staged
through pinned memory); CUDA torch tensors in -> CUDA torch tensor out (no host round trip).

Tie rule (stated with every result): fp32 squared distance ``((dx*dx)+(dy*dy))+(dz*dz)``, ``d = q - p``, no
FMA; ascending (distance, index).  On tie-free clouds this is bit-identical to nanoflann; on lattice clouds
nanoflann orders/keeps equal-distance points in kd-tree visiting order, so only the distance rows are
guaranteed identical.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib


class ConfigPancreas:
    """Hyper-parameters of the Pancreas model (``helper_tool.py:52-75``)."""
    k_n = 16
    num_layers = 5
    num_points = 180000
    num_classes = 2
    sub_grid_size = 0.01
    batch_size = 1
    val_batch_size = 1
    sub_sampling_ratio = [4, 4, 4, 4, 2]
    d_out = [16, 64, 128, 256, 512]
    num_features = 4  # xyz + 1 intensity channel (runPancreas.py:125)
    learning_rate = 1e-3
    lr_decays = 0.95
    name = "Pancreas"


class ConfigBraTS:
    """Hyper-parameters of the BraTS model (``helper_tool.py:21-51``); ``num_points`` follows BASELINE.json."""
    k_n = 16
    num_layers = 5
    num_points = 180000
    num_classes = 4
    sub_grid_size = 0.01
    batch_size = 4
    val_batch_size = 1
    sub_sampling_ratio = [4, 4, 4, 4, 2]
    d_out = [16, 64, 128, 256, 512]
    num_features = 7  # xyz + 4 MRI modalities (runBraTS.py:141)
    learning_rate = 1e-4
    lr_decays = 0.95
    name = "BraTS20"


_workspaces: dict[tuple[int, int], torch.Tensor] = {}


def workspace(nbytes: int, device: torch.device, slot: int = 0) -> torch.Tensor:
    """Grow-only scratch buffer per (device, slot); the C-ABI never allocates."""
    key = (device.index if device.index is not None else torch.cuda.current_device(), slot)
    buf = _workspaces.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(int(nbytes * 1.25) + 1024, dtype=torch.uint8, device=device)
        _workspaces[key] = buf
    return buf


def _stream_ptr(device) -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def knn_search_cuda(support: torch.Tensor, query: torch.Tensor, k: int, return_dist: bool = False,
                    out: torch.Tensor | None = None):
    """Device-resident entry: ``support [B,N1,3]``, ``query [B,N2,3]`` fp32 CUDA -> int32 ``[B,N2,k]`` CUDA.
    ``out`` (optional, contiguous int32 ``[B,N2,k]``) receives the result -- lets a caller allocate on one stream and
    search on another."""
    if not (support.is_cuda and query.is_cuda):
        raise _lib.PointUnetError("knn_search_cuda needs CUDA tensors (there is no CPU fallback)")
    if support.dim() != 3 or query.dim() != 3 or support.shape[2] != 3 or query.shape[2] != 3 \
            or support.shape[0] != query.shape[0]:
        raise ValueError(f"expected support [B,N1,3] and query [B,N2,3], got {tuple(support.shape)} / {tuple(query.shape)}")
    same = support.data_ptr() == query.data_ptr() and support.shape == query.shape
    support = support.contiguous().float()
    query = support if same else query.contiguous().float()
    B, N1, _ = support.shape
    N2 = query.shape[1]
    L = _lib.lib()
    if out is None:
        out = torch.empty((B, N2, k), dtype=torch.int32, device=support.device)
    elif out.dtype != torch.int32 or tuple(out.shape) != (B, N2, k) or not out.is_contiguous() or out.device != support.device:
        raise ValueError("knn_search_cuda: out must be a contiguous int32 [B,N2,k] tensor on the inputs' device")
    else:
        from . import ops  # the buffer is rewritten through its raw pointer: cached inverse lists of it are stale
        ops.drop_inverse(out.data_ptr())
    nbytes = L.pu_knn_workspace_bytes(B, N1, N2, k)
    ws = workspace(nbytes, support.device)
    with torch.cuda.device(support.device):
        if return_dist:
            dist = torch.empty((B, N2, k), dtype=torch.float32, device=support.device)
            st = L.pu_knn_batch_dist(support.data_ptr(), query.data_ptr(), B, N1, N2, k, out.data_ptr(),
                                     dist.data_ptr(), ws.data_ptr(), ws.numel(), _stream_ptr(support.device))
            _lib.check(st, "pu_knn_batch_dist")
            return out, dist
        st = L.pu_knn_batch(support.data_ptr(), query.data_ptr(), B, N1, N2, k, out.data_ptr(), ws.data_ptr(),
                            ws.numel(), _stream_ptr(support.device))
        _lib.check(st, "pu_knn_batch")
    return out


# Fraction of the rows of a level whose K neighbours held no sub-cloud point, as seen by the most recent finished
# pu_knn_self_interp call for that (N, n_sub): with a random prefix it is ~1 %; when the prefix is a spatial region (a volume
# listed organ first) it is most rows, and two separate searches are the faster way.  Both are exact, so this only picks speed.
_interp_unresolved: dict = {}
_interp_pending: dict = {}
FUSED_INTERP_MAX_UNRESOLVED = 0.125


def knn_self_interp_cuda(points: torch.Tensor, k: int, n_sub: int, out_neigh: torch.Tensor | None = None,
                         out_interp: torch.Tensor | None = None, adaptive: bool = True):
    """One pyramid level of ``tf_map`` (runPancreas.py:131-137): returns
    ``(knn_search(points, points, k) [B,N,k], knn_search(points[:, :n_sub], points, 1) [B,N,1])``, both int32.  By default from a
    single search structure (C-ABI ``pu_knn_self_interp``, bit-identical to the two searches); ``adaptive``: if the previous
    call for this shape found that most rows needed the filtered search (prefix = spatial region), run the two searches."""
    if not points.is_cuda:
        raise _lib.PointUnetError("knn_self_interp_cuda needs CUDA tensors (there is no CPU fallback)")
    if points.dim() != 3 or points.shape[2] != 3:
        raise ValueError(f"expected points [B,N,3], got {tuple(points.shape)}")
    points = points.contiguous().float()
    B, N, _ = points.shape
    dev = points.device
    from . import ops
    if out_neigh is None:
        out_neigh = torch.empty((B, N, k), dtype=torch.int32, device=dev)
    if out_interp is None:
        out_interp = torch.empty((B, N, 1), dtype=torch.int32, device=dev)
    for t, shape in ((out_neigh, (B, N, k)), (out_interp, (B, N, 1))):
        if t.dtype != torch.int32 or tuple(t.shape) != shape or not t.is_contiguous() or t.device != dev:
            raise ValueError("knn_self_interp_cuda: outputs must be contiguous int32 [B,N,k] / [B,N,1] tensors on the inputs' device")
        ops.drop_inverse(t.data_ptr())   # rewritten through the raw pointer: cached inverse lists of it are stale
    key = (dev.index, B, N, int(n_sub), int(k))
    pend = _interp_pending.get(key)
    if pend is not None and pend[1].query():   # the count of an earlier call has arrived on the host
        _interp_unresolved[key] = float(pend[0].sum()) / max(B * N, 1)
        _interp_pending.pop(key)
    if adaptive and _interp_unresolved.get(key, 0.0) > FUSED_INTERP_MAX_UNRESOLVED and n_sub >= 1:
        knn_search_cuda(points, points, k, out=out_neigh)
        knn_search_cuda(points[:, :n_sub].contiguous(), points, 1, out=out_interp)
        return out_neigh, out_interp
    L = _lib.lib()
    ws = workspace(L.pu_knn_workspace_bytes(B, N, N, k), dev)
    capturing = torch.cuda.is_current_stream_capturing()
    track = adaptive and not capturing and key not in _interp_pending and key not in _interp_unresolved
    cnt = torch.empty(B, dtype=torch.int32, device=dev) if track else None
    with torch.cuda.device(dev):
        st = L.pu_knn_self_interp(points.data_ptr(), B, N, int(k), int(n_sub), out_neigh.data_ptr(), out_interp.data_ptr(),
                                  cnt.data_ptr() if track else None, ws.data_ptr(), ws.numel(), _stream_ptr(dev))
        _lib.check(st, "pu_knn_self_interp")
    if track:   # fetched without a synchronisation; read by a later call once the copy has completed
        host = torch.empty(B, dtype=torch.int32, pin_memory=True)
        host.copy_(cnt, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(dev))
        _interp_pending[key] = (host, ev, cnt)
    return out_neigh, out_interp


def knn_last_stats(device=None) -> dict:
    """Counters of the last KNN call on ``device``: candidate distance evaluations, buckets swept, box tests."""
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    ws = workspace(0, device)
    arr = (ctypes.c_ulonglong * 3)()
    _lib.check(_lib.lib().pu_knn_read_stats(ws.data_ptr(), arr, _stream_ptr(device)), "pu_knn_read_stats")
    return dict(dist_evals=int(arr[0]), buckets=int(arr[1]), box_tests=int(arr[2]))


class DataProcessing:
    @staticmethod
    def knn_search(support_pts, query_pts, k):
        """
        :param support_pts: points you have, B*N1*3
        :param query_pts: points you want to know the neighbour index, B*N2*3
        :param k: Number of neighbours in knn search
        :return: neighbor_idx: neighboring points indexes, B*N2*k  (int32)

        Same contract as ``helper_tool.py:84-94``.  numpy arrays are staged host->device->host; CUDA
        tensors stay on the device.
        """
        if isinstance(support_pts, torch.Tensor) and support_pts.is_cuda:
            return knn_search_cuda(support_pts, query_pts, int(k))
        if not torch.cuda.is_available():
            raise _lib.PointUnetError("DataProcessing.knn_search needs a CUDA device (no CPU fallback)")
        same = support_pts is query_pts
        s = torch.from_numpy(np.ascontiguousarray(support_pts, dtype=np.float32)).cuda(non_blocking=True)
        q = s if same else torch.from_numpy(np.ascontiguousarray(query_pts, dtype=np.float32)).cuda(non_blocking=True)
        return knn_search_cuda(s, q, int(k)).cpu().numpy()

    @staticmethod
    def get_class_weights(dataset_name):
        """``helper_tool.py:172-184``: 1 / (class frequency + 0.02) with uniform pre-counted frequencies."""
        if dataset_name == "BraTS20":
            num_per_class = np.array([1, 1, 1, 1])
        elif dataset_name == "BraTS_Block64":
            num_per_class = np.array([1403, 22, 80, 11])
        elif dataset_name == "Pancreas":
            num_per_class = np.array([1, 1])
        else:
            raise ValueError(dataset_name)
        weight = num_per_class / float(sum(num_per_class))
        return np.expand_dims(1 / (weight + 0.02), axis=0)
