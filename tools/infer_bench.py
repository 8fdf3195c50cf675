"""Test-mode inference throughput (BASELINE.json configs[4] at single-GPU scale): per-volume point-cloud segmentation with
moving BN statistics + softmax + point2prod scatter into dense [Z,Y,X,C] volumes (testPancreas.py:141-202).
   python tools/infer_bench.py [--volumes 8] [--batch 4]  -> one JSON line (volumes/s, points/s, ms per batch)."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from point_unet_b200 import synthetic as syn
from point_unet_b200.helper_tool import ConfigPancreas
from point_unet_b200.train import Trainer

ap = argparse.ArgumentParser()
ap.add_argument("--volumes", type=int, default=8)
ap.add_argument("--batch", type=int, default=4)
ap.add_argument("--points", type=int, default=180000)
a = ap.parse_args()


class cfg(ConfigPancreas):
    num_points = a.points


shape = syn.PANCREAS_SHAPE                                   # (X, Y, Z) = (512, 512, 240)
vshape = (shape[2], shape[0], shape[1], 2)                   # the reference allocates (Z, X, Y, C)
clouds = [syn.pancreas_cloud(a.points, s) for s in range(a.batch)]
xyz = np.stack([c["xyz"] for c in clouds]); feats = np.stack([c["features"] for c in clouds])
xo = [c["xyz_origin"].astype(np.int32) for c in clouds]
tr = Trainer(cfg, num_features=4, seed=0, device="cuda")
pin = {k: torch.from_numpy(v).pin_memory() for k, v in dict(xyz=xyz, feats=feats).items()}
xo_dev = [torch.from_numpy(v).cuda() for v in xo]
for _ in range(2):
    vols = tr.predict_to_volume(pin["xyz"], pin["feats"], xo_dev, vshape)
torch.cuda.synchronize()
nb = max(a.volumes // a.batch, 1)
t0 = time.perf_counter()
for _ in range(nb):
    vols = tr.predict_to_volume(pin["xyz"], pin["feats"], xo_dev, vshape)   # H2D of the batch inside the timed region
    lab = [v.argmax(-1) for v in vols]                                       # genSegmentation*.py: arg-max over classes
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(json.dumps(dict(workload="test-mode inference + point2prod, Pancreas-shaped volumes 512x512x240, 180k points each",
                      volumes=nb * a.batch, batch=a.batch, ms_per_batch=dt / nb * 1e3, volumes_per_s=nb * a.batch / dt,
                      points_per_s=nb * a.batch * a.points / dt, gpu=torch.cuda.get_device_name(0))))
