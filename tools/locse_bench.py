"""Micro-benchmark of the fused position branch (csrc/locse_mlp.cu) per pyramid level: moments, forward, backward.
   python tools/locse_bench.py  -> one JSON line per (kernel, level); bytes = what the kernel has to move (idx + outputs /
   gradients), xyz is L2-resident."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from point_unet_b200 import ops
from point_unet_b200.helper_tool import knn_search_cuda, workspace
from point_unet_b200 import synthetic as syn
from bench import load_peaks

PEAK = load_peaks()["hbm"]
B, K = 4, 16
NL = [180000, 45000, 11250, 2812, 703]
DOUT = [16, 64, 128, 256, 512]


def timeit(fn, n=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    tot = 0.0
    for _ in range(n):
        flush.zero_()
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / n


def report(op, level, nbytes, ms, **kw):
    gbs = nbytes / (ms * 1e-3) / 1e9
    print(json.dumps(dict(op=op, level=level, ms=round(ms, 4), algorithmic_mb=round(nbytes / 1e6, 1), gbs=round(gbs, 1),
                          frac_of_measured_hbm=round(gbs / PEAK, 3), **kw)), flush=True)


xyz = torch.from_numpy(syn.batch(syn.brats_cloud, B, NL[0], 0)["xyz"]).cuda()
L = ops._L()
LEVELS = [int(v) for v in os.environ.get("LOCSE_LEVELS", "0,1,2,3,4").split(",")]
for lvl in LEVELS:
    N, h = NL[lvl], DOUT[lvl] // 2
    x3 = xyz[:, :N].contiguous()
    idx = knn_search_cuda(x3, x3, K)
    x = torch.empty(B, N, 4, device="cuda")
    ops._call("pu_locse_pack_xyz", x3.data_ptr(), B * N, x.data_ptr(), ops._stream(x3))
    R = B * N * K
    w = torch.randn(10, h, device="cuda") * 0.3
    bias = torch.zeros(h, device="cuda"); gamma = torch.ones(h, device="cuda"); beta = torch.zeros(h, device="cuda")
    mom = torch.empty(65, device="cuda"); coef = torch.empty(5 * h + 112, device="cuda")
    ws = workspace(L.pu_locse_mlp_workspace_bytes(h), x.device, slot=6)
    st = ops._stream(x)
    mo = lambda: ops._call("pu_locse_moments", x.data_ptr(), idx.data_ptr(), B, N, K, mom.data_ptr(), ws.data_ptr(), ws.numel(), st)
    ms = timeit(mo)
    report("locse_moments(2 passes)", lvl, 2 * 4 * R, ms, h=h)
    ops._call("pu_locse_bn_prepare", mom.data_ptr(), R, w.data_ptr(), h, bias.data_ptr(), gamma.data_ptr(), beta.data_ptr(), 1e-6, 1,
              None, None, 0.99, 1.0, coef.data_ptr(), st)
    buf = torch.empty(B, N, K, 2 * h, device="cuda"); fx = torch.empty(B, N, K, h, device="cuda")
    fw = lambda: ops._call("pu_locse_mlp_fwd", x.data_ptr(), idx.data_ptr(), B, N, K, w.data_ptr(), h, coef.data_ptr(), 0.2,
                           buf.data_ptr() + 4 * h, 2 * h, fx.data_ptr(), h, st)
    ms = timeit(fw)
    report("locse_mlp_fwd", lvl, 4 * R + 2 * 4 * R * h, ms, h=h)
    dbuf = torch.randn(B, N, K, 2 * h, device="cuda"); dfx = torch.randn(B, N, K, h, device="cuda")
    dw = torch.empty(10, h, device="cuda"); dg = torch.empty(h, device="cuda"); db = torch.empty(h, device="cuda")
    bw = lambda: ops._call("pu_locse_mlp_bwd", x.data_ptr(), idx.data_ptr(), B, N, K, w.data_ptr(), h, coef.data_ptr(), gamma.data_ptr(),
                           bias.data_ptr(), 1, 0.2, dbuf.data_ptr() + 4 * h, 2 * h, dfx.data_ptr(), h, dw.data_ptr(), 0, None,
                           dg.data_ptr(), db.data_ptr(), ws.data_ptr(), ws.numel(), st)
    ms = timeit(bw)
    report("locse_mlp_bwd", lvl, 4 * R + 2 * 4 * R * h, ms, h=h)
