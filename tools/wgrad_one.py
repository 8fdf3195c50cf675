"""One weight-gradient launch of a given shape (for ncu captures).  python tools/wgrad_one.py M K N"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from point_unet_b200 import ops
M, K, N = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (2880000, 64, 64)
x = torch.randn(M, K, device="cuda"); dy = torch.randn(M, N, device="cuda")
for _ in range(3): ops.wgrad_raw(x, dy)
torch.cuda.synchronize()
