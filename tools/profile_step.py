"""One training step bracketed by cudaProfilerStart/Stop, for `ncu --profile-from-start off ...`.
Usage: python tools/profile_step.py [--points N] [--batch B] [--mode train|knn]"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import make_batch
from point_unet_b200.helper_tool import ConfigBraTS, knn_search_cuda
from point_unet_b200.train import Trainer

ap = argparse.ArgumentParser()
ap.add_argument("--points", type=int, default=180000)
ap.add_argument("--batch", type=int, default=4)
ap.add_argument("--mode", default="train")
a = ap.parse_args()

class cfg(ConfigBraTS):
    num_points = a.points

host = make_batch(0, a.batch, a.points)
x = torch.from_numpy(host["xyz"]).cuda(); f = torch.from_numpy(host["features"]).cuda(); l = torch.from_numpy(host["labels"]).cuda()
if a.mode == "knn":
    for _ in range(2): knn_search_cuda(x, x, 16)
    torch.cuda.synchronize(); torch.cuda.profiler.start()
    knn_search_cuda(x, x, 16)
    torch.cuda.synchronize(); torch.cuda.profiler.stop()
else:
    tr = Trainer(cfg, num_features=7, device="cuda")
    for _ in range(2): tr.train_step_device(x, f, l)
    torch.cuda.synchronize(); torch.cuda.profiler.start()
    tr.train_step_device(x, f, l)
    torch.cuda.synchronize(); torch.cuda.profiler.stop()
print("profiled one step")
