import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from point_unet_b200 import ops
from point_unet_b200.helper_tool import workspace
torch.manual_seed(0)
M,K,N = 4096,64,64
x = torch.randn(M,K,device="cuda"); dy = torch.randn(M,N,device="cuda")
want = x.double().t() @ dy.double()
L = ops._L()
nbytes = L.pu_tc_wgrad_workspace_bytes(M,K,N)
ws = workspace(nbytes, x.device, slot=2)
ws.view(torch.float32)[: nbytes//4].fill_(float("nan"))
ops.tc_error_flag(x.device).zero_()
dw,_ = ops.wgrad_raw(x, dy, tc_mode=1)
torch.cuda.synchronize()
part = ws.view(torch.float32)[: nbytes//4]
print("nbytes", nbytes, "flag", int(ops.tc_error_flag(x.device).item()))
print("part nan frac", float(torch.isnan(part).float().mean()), "zero frac", float((part==0).float().mean()))
p0 = part[:K*N].view(K,N)
print("cta0 partial[0,:4]", p0[0,:4].tolist(), " expected (first 128 rows):", (x[:128].double().t() @ dy[:128].double())[0,:4].tolist())
print("dw", dw[0,:4].tolist())
