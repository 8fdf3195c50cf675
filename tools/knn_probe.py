"""Quick KNN timing probe (CUDA events, device-resident inputs).  Usage: python tools/knn_probe.py"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from point_unet_b200 import synthetic as syn
from point_unet_b200.helper_tool import knn_search_cuda, knn_last_stats

def timeit(fn, n=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

for kind, n in [("uniform", 16384), ("uniform", 65536), ("uniform", 180000), ("pancreas", 180000), ("brats", 180000), ("uniform", 1000000)]:
    if kind == "uniform": p = syn.uniform_cloud(n)
    elif kind == "pancreas": p = syn.pancreas_cloud(n, 0)["xyz"]
    else: p = syn.brats_cloud(n, 0)["xyz"]
    for B in (1, 4):
        if B == 4 and n > 180000: continue
        t = torch.from_numpy(np.stack([p] * B)).cuda()
        ms = timeit(lambda: knn_search_cuda(t, t, 16))
        st = knn_last_stats()
        sub = t[:, : n // 4].contiguous()
        ms1 = timeit(lambda: knn_search_cuda(sub, t, 1))
        st1 = knn_last_stats()
        print(f"{kind:9s} N={n:8d} B={B} K16 self: {ms:8.3f} ms {B*n/ms/1e3:9.1f} Mq/s evals/q {st['dist_evals']/(B*n):7.1f} buckets/warp {st['buckets']/(B*n/32):6.1f} tests/warp {st['box_tests']/(B*n/32):7.1f}"
              f" | K1 prefix: {ms1:7.3f} ms {B*n/ms1/1e3:9.1f} Mq/s evals/q {st1['dist_evals']/(B*n):6.1f}", flush=True)
