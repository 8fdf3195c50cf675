import sys, os, time
sys.path.insert(0, "/root/repo")
import torch, numpy as np
from point_unet_b200 import synthetic as syn
from point_unet_b200.helper_tool import knn_search_cuda, knn_self_interp_cuda
def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for name, gen in (("brats", syn.brats_cloud), ("pancreas", syn.pancreas_cloud)):
    for N in (180000, 45000, 11250):
        x = torch.from_numpy(syn.batch(gen, 4, 180000, 0)["xyz"]).cuda()[:, :N].contiguous()
        ns = N // 4
        a = t(lambda: knn_self_interp_cuda(x, 16, ns))
        b = t(lambda: (knn_search_cuda(x, x, 16), knn_search_cuda(x[:, :ns].contiguous(), x, 1)))
        nb, it = knn_self_interp_cuda(x, 16, ns)
        unresolved = int(((nb >= ns).all(dim=-1)).sum())
        print(name, N, "fused ms", round(a, 3), "separate ms", round(b, 3), "unresolved rows", unresolved, flush=True)
