"""Per-shape timing of the tcgen05 linear kernel (forward with statistics, dgrad, accumulate-dgrad of att_pooling) over the
layer table of the benchmark config (B = 4 x 180k), L2 flushed between runs.   python tools/linear_bench.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from point_unet_b200 import ops
from bench import load_peaks
PEAK = load_peaks()["hbm"]
B, K = 4, 16
NL = [180000, 45000, 11250, 2812, 703, 351]
DOUT = [16, 64, 128, 256, 512]
shapes = []   # (name, M, Kin, N, accumulate, stats)
d_in = 8
for i, d in enumerate(DOUT):
    n = B * NL[i]
    for nm, M, a, b in (("mlp1", n, d_in, d // 2), ("att1mlp", n, d, d // 2), ("LFAmlp2", n * K, d // 2, d // 2), ("att2mlp", n, d, d),
                        ("mlp2", n, d, 2 * d), ("shortcut", n, d_in, 2 * d)):
        shapes.append(("L%d %s fwd" % (i, nm), M, a, b, 0, 1))
        shapes.append(("L%d %s dgrad" % (i, nm), M, b, a, 0, 0))
    shapes.append(("L%d att dgrad-acc (x2)" % i, n * K, d, d, 1, 0))
    d_in = 2 * d
feat, enc = 1024, [32, 32, 128, 256, 512, 1024]
shapes.append(("decoder_0 fwd", B * NL[5], 1024, 1024, 0, 1)); shapes.append(("decoder_0 dgrad", B * NL[5], 1024, 1024, 0, 0))
for j in range(5):
    skip = enc[-j - 2]
    shapes.append(("Decoder_%d fwd" % j, B * NL[4 - j], skip + feat, skip, 0, 1))
    shapes.append(("Decoder_%d dgrad" % j, B * NL[4 - j], skip, skip + feat, 0, 0))
    feat = skip
shapes += [("fc1 fwd", B * NL[0], 32, 64, 0, 1), ("fc1 dgrad", B * NL[0], 64, 32, 0, 0), ("fc2 fwd", B * NL[0], 64, 32, 0, 1), ("fc2 dgrad", B * NL[0], 32, 64, 0, 0)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
tot = 0.0
for name, M, Kin, N, acc, stats in shapes:
    if Kin < 32 or N < 32:
        continue
    x = torch.randn(M, Kin, device="cuda"); wt = (torch.randn(N, Kin, device="cuda") / Kin ** 0.5).contiguous()
    out = torch.zeros(M, N, device="cuda")
    fn = lambda: ops.linear_raw(x, None, None, out=out, accumulate=bool(acc), want_stats=bool(stats), wt=wt)
    for _ in range(3): fn()
    ms = 0.0
    for _ in range(5):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ms += e0.elapsed_time(e1) / 5
    nbytes = 4 * M * (Kin + N * (1 + acc))
    mult = 2 if "x2" in name else 1
    tot += ms * mult
    print(json.dumps(dict(layer=name, M=M, K=Kin, N=N, acc=acc, stats=stats, ms=round(ms, 4), mb=round(nbytes / 1e6, 1),
                          gbs=round(nbytes / ms / 1e6, 0), frac=round(nbytes / ms / 1e6 / PEAK, 3))), flush=True)
print(json.dumps(dict(total_ms=round(tot, 3), note="want_stats launches include the stats_finalize kernel")))
