"""Runs a few tensor-core linear shapes (for ncu): python tools/tc_probe.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from point_unet_b200 import ops
torch.manual_seed(0)
M = 2880000
def run(K, N, acc, stats):
    x = torch.randn(M, K, device="cuda"); w = torch.randn(K, N, device="cuda") * 0.1
    out = torch.randn(M, N, device="cuda")
    for _ in range(2): ops.linear_raw(x, w, None, out=out, accumulate=acc, want_stats=stats)
    torch.cuda.synchronize(); torch.cuda.profiler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.linear_raw(x, w, None, out=out, accumulate=acc, want_stats=stats); e1.record()
    torch.cuda.synchronize(); torch.cuda.profiler.stop()
    ms = e0.elapsed_time(e1); gb = 4 * M * (K + N * (2 if acc else 1)) / 1e9
    print(f"K={K} N={N} acc={acc} stats={stats}: {ms:.3f} ms  {gb/ms:.2f} TB/s algorithmic", flush=True)
run(32, 32, False, False)
run(64, 64, True, False)
run(32, 32, False, True)
