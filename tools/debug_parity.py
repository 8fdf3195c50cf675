"""Per-tensor parity table of the full network vs the fp64 / fp32 oracle (debugging aid).  python tools/debug_parity.py [N] [B]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import randla_ref as ref
from point_unet_b200 import synthetic as syn
from point_unet_b200.helper_tool import ConfigPancreas
from point_unet_b200.RandLANet import Network, build_pyramid, init_params, layer_table

n_points = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2
class cfg(ConfigPancreas): num_points = n_points
data = syn.batch(syn.pancreas_cloud, B, n_points, seed0=40)
params = init_params(cfg, 4, seed=1)
rng = np.random.default_rng(2)
for k in params:
    if k.endswith("gamma"): params[k] = (params[k] + rng.uniform(-0.3, 0.3, params[k].shape)).astype(np.float32)
    if k.endswith("beta") or k.endswith("biases") or k.endswith("bias"): params[k] = rng.uniform(-0.1, 0.1, params[k].shape).astype(np.float32)
net = Network(cfg, 4, device="cuda"); net.load_numpy(params)
xyz = torch.from_numpy(data["xyz"]).cuda()
pyr = build_pyramid(xyz, cfg)
feats = torch.cat([xyz, torch.from_numpy(data["features"]).cuda()], dim=-1)
labels = torch.from_numpy(data["labels"]).cuda()
mask = torch.from_numpy(rng.random((B, n_points, 1, 32)) < 0.5).cuda()
logits = net.inference(dict(pyr, features=feats), True, dropout_mask=mask)
loss = net.get_loss(logits, labels); loss.backward()
def run(dt):
    pp = {k: torch.from_numpy(v).to(dt).requires_grad_("moving" not in k) for k, v in params.items()}
    inp = dict(xyz=[t.cpu().to(dt) for t in pyr["xyz"]], neigh_idx=[t.cpu() for t in pyr["neigh_idx"]], sub_idx=[t.cpu() for t in pyr["sub_idx"]],
               interp_idx=[t.cpu() for t in pyr["interp_idx"]], features=feats.cpu().to(dt))
    lg = ref.inference(pp, inp, cfg, True, dropout_mask=mask.cpu()); ls = ref.get_loss(lg, labels.cpu(), net.class_weights.cpu().numpy()); ls.backward()
    return pp, lg
p64, l64 = run(torch.float64); p32, l32 = run(torch.float32)
l2 = lambda a, b: float((a.double().cpu() - b.double().cpu()).norm() / max(float(b.double().norm()), 1e-30))
print("logits ours %.2e fp32 %.2e" % (l2(logits.detach(), l64.detach()), l2(l32.detach(), l64.detach())))
for scope, kind, cin, cout in layer_table(cfg, 4):
    for suffix in ("/kernel", "/weights", "/bn/gamma", "/bn/beta"):
        name = scope + suffix
        if name in p64 and p64[name].grad is not None:
            g = dict(net.named_variables())[name].grad
            print("%-42s %-10s ours %.2e  fp32 %.2e  |g| %.2e" % (name, str(tuple(g.shape)), l2(g, p64[name].grad), l2(p32[name].grad, p64[name].grad), float(p64[name].grad.norm())))
