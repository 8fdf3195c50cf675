import sys, os, time
sys.path.insert(0, "/root/repo")
import torch
from point_unet_b200 import ops
P = int(sys.argv[1]); K, d = 16, 64
g = torch.Generator().manual_seed(P)
x = torch.randn(1, P, K, d, generator=g).cuda().requires_grad_(True)
w = (torch.randn(d, d, generator=g) * 0.2).cuda().requires_grad_(True)
dy = torch.randn(1, P, 1, d, generator=g).cuda()
ops.tc_error_flag(x.device).zero_()
agg = ops.att_pool(x, w)
torch.cuda.synchronize(); print("fwd ok", flush=True)
t0 = time.time()
(agg * dy).sum().backward()
torch.cuda.synchronize(); print("bwd done in", time.time() - t0, "flag", int(ops.tc_error_flag(x.device).item()), flush=True)
xr, wr = x.detach().cpu().double().requires_grad_(True), w.detach().cpu().double().requires_grad_(True)
act = xr.reshape(-1, K, d) @ wr
(( xr.reshape(-1, K, d) * torch.softmax(act, dim=1)).sum(1).reshape(1, P, 1, d) * dy.cpu().double()).sum().backward()
print("dx err", float((x.grad.cpu().double() - xr.grad).abs().max() / xr.grad.abs().max()), "dw err", float((w.grad.cpu().double() - wr.grad).abs().max() / wr.grad.abs().max()))
