"""Times the tiny per-channel kernels (batch-norm coefficient / statistics reductions) on the shapes of one training step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes, torch
from point_unet_b200 import ops
L = ops._L()
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
def timeit(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for blocks, C in ((1184, 16), (1184, 64), (1184, 8), (1184, 128), (352, 512), (88, 1024)):
    p1 = torch.randn(blocks, C, device="cuda"); p2 = torch.randn(blocks, C, device="cuda")
    mean = torch.randn(C, device="cuda"); inv = torch.rand(C, device="cuda") + 0.5; gamma = torch.rand(C, device="cuda")
    co = torch.empty(5, C, device="cuda")
    t = timeit(lambda: L.pu_bn_bwd_coeffs(p1.data_ptr(), p2.data_ptr(), blocks, C, mean.data_ptr(), inv.data_ptr(), gamma.data_ptr(), 1000000, 1,
                                          co[0].data_ptr(), co[1].data_ptr(), co[2].data_ptr(), co[3].data_ptr(), co[4].data_ptr(), st))
    print(f"bn_bwd_coeffs blocks={blocks} C={C}: {t:.1f} us")
for tiles, rpt, C in ((5625, 2048, 16), (22500, 128, 64), (22500, 128, 32), (5625, 128, 128), (1406, 128, 256), (352, 128, 512)):
    s = torch.randn(tiles, C, device="cuda"); m2 = torch.rand(tiles, C, device="cuda")
    mean = torch.empty(C, device="cuda"); var = torch.empty(C, device="cuda")
    t = timeit(lambda: L.pu_stats_finalize(s.data_ptr(), m2.data_ptr(), tiles, rpt, C, tiles * rpt, mean.data_ptr(), var.data_ptr(), st))
    print(f"stats_finalize tiles={tiles} C={C}: {t:.1f} us")
