"""Per-role clock64 timeline of tc_wgrad_ts_kernel (CTA 0, first 96 steps).  python tools/wgrad_timeline.py M K N"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from point_unet_b200 import _lib, ops
M, K, N = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (2880000, 64, 64)
L = _lib.lib()
r, s, e = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
L.pu_tc_debug_wgrad_timeline_dims(ctypes.byref(r), ctypes.byref(s), ctypes.byref(e))
R, S, E = r.value, s.value, e.value
x = torch.randn(M, K, device="cuda"); dy = torch.randn(M, N, device="cuda")
for _ in range(2): ops.wgrad_raw(x, dy)
buf = torch.zeros(R * S * E, dtype=torch.int64, device="cuda")
L.pu_tc_debug_set_wgrad_timeline(ctypes.c_void_p(buf.data_ptr()))
ops.wgrad_raw(x, dy)
torch.cuda.synchronize()
L.pu_tc_debug_set_wgrad_timeline(None)
t = buf.view(R, S, E).cpu()
t0 = int(t[t > 0].min())
names = ["A-conv: raw_full | a_free | st issued | a_ready arrived", "B-conv: raw_full | b_free | STS done | b_ready arrived",
         "issuer: a_ready | b_ready | MMAs issued | committed", "loader: raw_free (slot reusable) | - | - | -"]
print(f"M={M} K={K} N={N}; cycles relative to the first stamp; steps 16..47")
for role in range(R):
    print(names[role])
    for step in range(16, 48):
        print(f"  step {step:3d}: " + " ".join(f"{int(v) - t0:8d}" if v > 0 else "       -" for v in t[role, step]))
per = (int(t[2, 80, 3]) - int(t[2, 16, 3])) / 64.0
print(f"issuer period over steps 16..80: {per:.0f} cycles per step")
