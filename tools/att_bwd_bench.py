"""Timing of the d = 64 attentive-pooling backward (2.88 M rows): fused kernel vs two-kernel path.  python tools/att_bwd_bench.py [dbg]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from point_unet_b200 import ops
P, K, d = 180000, 16, 64
x = torch.randn(1, P, K, d, device="cuda"); w = (torch.randn(d, d, device="cuda") * 0.2); wt = w.t().contiguous()
g = torch.randn(P, d, device="cuda")
d_act = torch.empty(P * K, d, device="cuda"); dx = torch.empty(P * K, d, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
L = ops._L()
def timeit(fn, n=5):
    for _ in range(2): fn()
    ms = 0.0
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ms += e0.elapsed_time(e1) / n
    return ms
flag = ops.tc_error_flag(x.device)
tws = ops.workspace(L.pu_tc_workspace_bytes(d, d), x.device, slot=4)
def unfused():
    ops._call("pu_tc_att_pooling_bwd", x.data_ptr(), d, wt.data_ptr(), g.data_ptr(), d, P, K, d, d_act.data_ptr(), d, dx.data_ptr(), d, 3,
              flag.data_ptr(), tws.data_ptr(), tws.numel(), ops._stream(x))
def acc():
    ops.linear_raw(d_act, None, wt=w, out=dx, accumulate=True)
print("att bwd (two outputs)", round(timeit(unfused), 4), "ms; accumulate GEMM", round(timeit(acc), 4), "ms")
for dbg in [0]:
    def fused():
        ops._call("pu_tc_att_pooling_bwd_fused", x.data_ptr(), d, wt.data_ptr(), w.data_ptr(), g.data_ptr(), d, P, K, d, d_act.data_ptr(), d,
                  dx.data_ptr(), d, 3 | (dbg << 8), flag.data_ptr(), ops._stream(x))
    print("fused dbg", dbg, round(timeit(fused), 4), "ms", "flag", int(flag.item()))
