"""Micro-benchmark of the HBM-bound LFA kernels: achieved GB/s on SURVEY 8(d) algorithmic bytes vs the measured copy peak.
   python tools/op_bench.py  -> one JSON line per (op, level)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from point_unet_b200 import ops
from point_unet_b200.helper_tool import knn_search_cuda
from point_unet_b200 import synthetic as syn
from bench import load_peaks

PEAK = load_peaks()["hbm"]
B, K = 4, 16
NL = [180000, 45000, 11250, 2812, 703, 351]
DOUT = [16, 64, 128, 256, 512]

def timeit(fn, n=20, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    tot = 0.0
    for _ in range(n):
        flush.zero_()  # flush L2 between iterations (buffer > 126 MB)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / n

def report(op, level, nbytes, ms, **kw):
    gbs = nbytes / (ms * 1e-3) / 1e9
    print(json.dumps(dict(op=op, level=level, ms=round(ms, 4), algorithmic_mb=round(nbytes / 1e6, 1), gbs=round(gbs, 1),
                          frac_of_measured_hbm=round(gbs / PEAK, 3), **kw)), flush=True)

import numpy as np
def locality_order(xyz_np):
    """(class, Morton) order: class = pyramid level at which a point drops out (prefix sets are preserved)."""
    n = xyz_np.shape[0]
    cls = np.zeros(n, np.int64)
    for lvl, nl in enumerate(NL[:-1]):   # points >= NL[lvl+1] and < NL[lvl] drop out after level lvl
        cls[NL[lvl + 1]:nl] = len(NL) - 1 - lvl
    lo, hi = xyz_np.min(0), xyz_np.max(0)
    q = np.clip(((xyz_np - lo) / np.maximum(hi - lo, 1e-12) * 512).astype(np.int64), 0, 511)
    def spread(v):
        r = np.zeros_like(v)
        for b in range(9):
            r |= ((v >> b) & 1) << (3 * b)
        return r
    m = (spread(q[:, 0]) << 2) | (spread(q[:, 1]) << 1) | spread(q[:, 2])
    return np.argsort((cls << 27) | m, kind="stable")
data = syn.batch(syn.brats_cloud, B, NL[0], 0)["xyz"]
if "--reorder" in sys.argv:
    data = np.stack([c[locality_order(c)] for c in data])
xyz = torch.from_numpy(data).cuda()
for lvl in range(3):
    N, Nn, d = NL[lvl], NL[lvl + 1], DOUT[lvl]
    x = xyz[:, :N].contiguous()
    idx = knn_search_cuda(x, x, K)
    h = d // 2
    f = torch.randn(B, N, h, device="cuda")
    out = torch.empty(B, N, K, h, device="cuda")
    ms = timeit(lambda: ops.gather_rows(f, idx, out=out))
    report("gather_neighbour", lvl, B * (4 * N * K + 4 * N * h + 4 * N * K * h), ms, d=h)
    inv = ops.inverse_of(idx, N)
    g = torch.randn(B, N, K, h, device="cuda")
    gs = torch.empty(B * N, h, device="cuda")
    ms = timeit(lambda: ops.segment_sum(g, inv, h, out=gs))
    report("gather_neighbour_bwd(segment_sum)", lvl, B * (4 * N * K + 4 * N * h + 4 * N * K * h), ms, d=h)
    ms = timeit(lambda: ops.relative_pos_encoding(x, idx))
    report("relative_pos_encoding", lvl, B * (4 * N * K + 12 * N + 40 * N * K), ms)
    feat = torch.randn(B, N, 1, 2 * d, device="cuda", requires_grad=True)
    pool = idx[:, :Nn].contiguous()
    ms = timeit(lambda: ops.random_sample(feat.detach(), pool))
    report("random_sample", lvl, B * (4 * Nn * K + 4 * N * 2 * d + 4 * Nn * 2 * d), ms, d=2 * d)
    o = ops.random_sample(feat, pool); go = torch.randn_like(o)
    ms = timeit(lambda: torch.autograd.grad(o, feat, go, retain_graph=True))
    report("random_sample_bwd", lvl, B * (4 * Nn * K + 2 * 4 * N * 2 * d + 2 * 4 * Nn * 2 * d), ms, d=2 * d)
    sub = x[:, :Nn].contiguous()
    up = knn_search_cuda(sub, x, 1)
    c = [32, 128, 256][lvl]
    fl = torch.randn(B, Nn, 1, c, device="cuda")
    ms = timeit(lambda: ops.nearest_interpolation(fl, up))
    report("nearest_interpolation", lvl, B * (4 * N + 4 * Nn * c + 4 * N * c), ms, d=c)
    y = torch.randn(B, N, K, h, device="cuda"); sc = torch.rand(h, device="cuda"); sh = torch.randn(2, h, device="cuda")
    ms = timeit(lambda: ops._bn_act_fwd_raw(y, sc, sh, 0.2, out=out))
    report("bn_act_fwd", lvl, 2 * 4 * B * N * K * h, ms, d=h)
    ops.clear_caches()
