"""att_pooling forward/backward and the linear kernel on the level-0/1 shapes, 3xTF32 vs single TF32 (a probe for what
bounds the tensor-core kernels: if halving the shared-memory operand traffic shortens them, they are smem-bound)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from point_unet_b200 import ops
torch.manual_seed(0)

def timeit(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

for (P, d) in ((180000, 64), (720000, 16), (45000, 128), (720000 // 4, 32)):
    x = torch.randn(1, P, 16, d, device="cuda", requires_grad=True); w = (torch.randn(d, d, device="cuda") * 0.2).requires_grad_()
    g = torch.randn(1, P, 1, d, device="cuda")
    for mode in (3, 1):
        ops.TC_MODE = mode
        with torch.no_grad():
            f = timeit(lambda: ops.att_pool(x, w))
        def fb():
            out = ops.att_pool(x, w); out.backward(g); x.grad = None; w.grad = None
        t = timeit(fb)
        gb = 4 * P * 16 * d / 1e9
        print(f"att P={P} d={d} mode={mode}: fwd {f:.3f} ms ({gb / f:.2f} TB/s), fwd+bwd {t:.3f} ms", flush=True)
M = 2880000
for (K, N, acc) in ((64, 64, False), (64, 64, True), (32, 32, False), (128, 128, False)):
    Mx = M if K <= 64 else M // 4
    x = torch.randn(Mx, K, device="cuda"); w = torch.randn(K, N, device="cuda") * 0.1; out = torch.randn(Mx, N, device="cuda")
    for mode in (3, 1):
        t = timeit(lambda: ops.linear_raw(x, w, None, out=out, accumulate=acc, want_stats=True, tc_mode=mode))
        gb = 4 * Mx * (K + N * (2 if acc else 1)) / 1e9
        print(f"linear M={Mx} K={K} N={N} acc={acc} mode={mode}: {t:.3f} ms ({gb / t:.2f} TB/s)", flush=True)
    dy = torch.randn(Mx, N, device="cuda")
    for mode in (3, 1):
        t = timeit(lambda: ops.wgrad_raw(x, dy, tc_mode=mode))
        print(f"wgrad  M={Mx} K={K} N={N} mode={mode}: {t:.3f} ms ({4 * Mx * (K + N) / 1e9 / t:.2f} TB/s)", flush=True)
