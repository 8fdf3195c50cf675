"""Per C-ABI entry point and per shape tag: launches and CUDA-event time of ONE training step (after warm-up).
Usage: python tools/step_breakdown.py [--points N] [--batch B] [--top T]"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import make_batch
from point_unet_b200 import _lib, ops
from point_unet_b200.helper_tool import ConfigBraTS
from point_unet_b200.train import Trainer

ap = argparse.ArgumentParser()
ap.add_argument("--points", type=int, default=180000)
ap.add_argument("--batch", type=int, default=4)
ap.add_argument("--top", type=int, default=70)
a = ap.parse_args()

class cfg(ConfigBraTS):
    num_points = a.points

host = make_batch(0, a.batch, a.points)
x = torch.from_numpy(host["xyz"]).cuda(); f = torch.from_numpy(host["features"]).cuda(); l = torch.from_numpy(host["labels"]).cuda()
tr = Trainer(cfg, num_features=7, device="cuda")
for _ in range(3):
    tr.train_step_device(x, f, l)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    tr.train_step_device(x, f, l)
e1.record(); torch.cuda.synchronize()
print(f"step (untimed entries): {e0.elapsed_time(e1) / 5:.3f} ms")
with ops.KernelTimer(set(_lib._OP_SIGS.keys())) as kt:
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record(); tr.train_step_device(x, f, l); s1.record()
    summ = kt.summary()
print(f"step (every entry bracketed by events): {s0.elapsed_time(s1):.3f} ms")
tot = sum(v[1] for v in summ.values())
print(f"sum over C-ABI entries: {tot:.3f} ms (the rest is torch glue: adds, fills, cats, optimizer)")
rows = []
for name, (n, ms, tags) in summ.items():
    for tag, (cnt, tms) in tags.items():
        rows.append((tms, name, tag, cnt))
print(f"{'ms':>8} {'n':>4}  entry / tag")
for name, (n, ms, tags) in sorted(summ.items(), key=lambda kv: -kv[1][1]):
    print(f"{ms:8.3f} {n:4d}  {name}")
print("---- by shape")
for tms, name, tag, cnt in sorted(rows, reverse=True)[:a.top]:
    print(f"{tms:8.3f} {cnt:4d}  {name} {tag}")
