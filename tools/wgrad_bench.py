"""Per-shape timing of the weight-gradient kernels over the layer table of the benchmark config (B = 4 x 180k).
   python tools/wgrad_bench.py   (PU_WGRAD_TS=0 for the first-generation kernel)"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from point_unet_b200 import ops
from bench import load_peaks
PEAK = load_peaks()["hbm"]
B, K = 4, 16
NL = [180000, 45000, 11250, 2812, 703, 351]
DOUT = [16, 64, 128, 256, 512]
shapes = []
d_in = 8
for i, d in enumerate(DOUT):
    n = B * NL[i]
    shapes += [("L%d mlp1" % i, n, d_in, d // 2), ("L%d LFAmlp1" % i, n * K, 10, d // 2), ("L%d att1fc" % i, n * K, d, d),
               ("L%d att1mlp" % i, n, d, d // 2), ("L%d LFAmlp2" % i, n * K, d // 2, d // 2), ("L%d att2fc" % i, n * K, d, d),
               ("L%d att2mlp" % i, n, d, d), ("L%d mlp2" % i, n, d, 2 * d), ("L%d shortcut" % i, n, d_in, 2 * d)]
    d_in = 2 * d
shapes.append(("decoder_0", B * NL[5], 1024, 1024))
feat, enc = 1024, [32, 32, 128, 256, 512, 1024]
for j in range(5):
    skip = enc[-j - 2]
    shapes.append(("Decoder_%d" % j, B * NL[4 - j], skip + feat, skip))
    feat = skip
shapes += [("fc1", B * NL[0], 32, 64), ("fc2", B * NL[0], 64, 32)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
tot = 0.0
for name, M, Kin, N in shapes:
    x = torch.randn(M, Kin, device="cuda"); dy = torch.randn(M, N, device="cuda")
    dw = torch.zeros(Kin, N, device="cuda")
    for _ in range(3): ops.wgrad_raw(x, dy, out=dw)
    ms = 0.0
    for _ in range(5):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.wgrad_raw(x, dy, out=dw); e1.record(); torch.cuda.synchronize()
        ms += e0.elapsed_time(e1) / 5
    nbytes = 4 * M * (Kin + N)
    tot += ms
    print(json.dumps(dict(layer=name, M=M, K=Kin, N=N, ms=round(ms, 4), mb=round(nbytes / 1e6, 1), gbs=round(nbytes / ms / 1e6, 0),
                          frac=round(nbytes / ms / 1e6 / PEAK, 3))), flush=True)
print(json.dumps(dict(total_ms=round(tot, 3))))
