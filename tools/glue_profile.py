"""torch.profiler table of one training step: which torch (non C-ABI) ops run, how often, and with which stacks."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from bench import make_batch
from point_unet_b200.helper_tool import ConfigBraTS
from point_unet_b200.train import Trainer

class cfg(ConfigBraTS):
    num_points = 180000
host = make_batch(0, 4, 180000)
x = torch.from_numpy(host["xyz"]).cuda(); f = torch.from_numpy(host["features"]).cuda(); l = torch.from_numpy(host["labels"]).cuda()
tr = Trainer(cfg, num_features=7, device="cuda")
for _ in range(3):
    tr.train_step_device(x, f, l)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True) as prof:
    tr.train_step_device(x, f, l)
    torch.cuda.synchronize()
ka = prof.key_averages()
tot = sum(e.self_device_time_total for e in ka)
print(f"total device time {tot / 1e3:.3f} ms")
print("---- device kernels that are not ours")
glue = 0.0
for e in sorted(ka, key=lambda e: -e.self_device_time_total):
    if e.self_device_time_total <= 0 or "pu::" in e.key:
        continue
    glue += e.self_device_time_total
    print(f"{e.self_device_time_total / 1e3:8.3f} ms {e.count:5d}  {e.key[:110]}")
print(f"glue total {glue / 1e3:.3f} ms")
print("---- aten ops by python stack (device time incl. children)")
ks = prof.key_averages(group_by_stack_n=8)
for e in sorted(ks, key=lambda e: -e.device_time_total)[:400]:
    if not e.key.startswith("aten::") or e.device_time_total < 20:
        continue
    st = [s for s in e.stack if "point_unet_b200" in s or "bench" in s][:3]
    print(f"{e.device_time_total / 1e3:8.3f} ms {e.count:5d}  {e.key:28s} {' <- '.join(x.split('point_unet_b200/')[-1] for x in st)}")
