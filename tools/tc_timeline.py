"""Per-role timeline of ONE launch of the persistent tcgen05 kernel (development tool, not part of the product path).

Builds the library variant with -DPU_TC_TIMELINE (`make TL=1` -> point_unet_b200/libpointunet_b200_tl.so), runs
pu_tc_linear_fwd once on the given shape and prints, for CTA (0,0), the clock64 stamps of the hand-off points of each warp
role per work item (k-block for loader / converter / issuer, (tile, pass) for the two epilogue groups):

  loader     e0 slot free, TMA issued
  converter  e0 x k-block landed   e1 TMEM stage free   e2 tcgen05.st issued   e3 stores complete, stage handed over
  issuer     e0 stage ready        e1 weight k-block landed   e2 MMAs issued   e3 next weight fetch issued
  epilogue   e0 pass start         e1 accumulator ready       e2 TMEM drained  e3 pass done

Usage: python tools/tc_timeline.py [--M 45000 --K 256 --N 256] [--accumulate] [--items 48]
"""
import argparse, ctypes, os, subprocess, sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

ap = argparse.ArgumentParser()
ap.add_argument("--M", type=int, default=45000)
ap.add_argument("--K", type=int, default=256)
ap.add_argument("--N", type=int, default=256)
ap.add_argument("--accumulate", action="store_true")
ap.add_argument("--items", type=int, default=48)
ap.add_argument("--mode", type=int, default=3)
a = ap.parse_args()

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "point_unet_b200", "csrc")
if not (os.environ.get("PU_TL_NOBUILD") and os.path.exists(os.path.join(ROOT, "point_unet_b200", "libpointunet_b200_tl.so"))):
    subprocess.check_call(["make", "-s", "-C", CSRC, "TL=1"], stderr=subprocess.DEVNULL)
L = ctypes.CDLL(os.path.join(ROOT, "point_unet_b200", "libpointunet_b200_tl.so"))
c_void_p, c_int, c_ll, c_size_t = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_size_t
L.pu_tc_workspace_bytes.restype = c_size_t
L.pu_tc_workspace_bytes.argtypes = [c_int, c_int]
L.pu_tc_linear_fwd.argtypes = [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int, c_ll, c_int, c_int, c_int, c_void_p,
                               c_void_p, c_int, c_void_p, c_void_p, c_size_t, c_void_p]
L.pu_tc_debug_set_timeline.argtypes = [c_void_p]
r, it, ev = c_int(), c_int(), c_int()
L.pu_tc_debug_timeline_dims(ctypes.byref(r), ctypes.byref(it), ctypes.byref(ev))
R, I, E = r.value, it.value, ev.value

dev = torch.device("cuda", 0)
x = torch.randn(a.M, a.K, device=dev)
wt = torch.randn(a.N, a.K, device=dev) / a.K ** 0.5
y = torch.zeros(a.M, a.N, device=dev)
ws = torch.empty(max(int(L.pu_tc_workspace_bytes(a.K, a.N)), 16), dtype=torch.uint8, device=dev)
flag = torch.zeros(1, dtype=torch.int32, device=dev)
tl = torch.zeros(R * I * E, dtype=torch.int64, device=dev)
st = c_void_p(torch.cuda.current_stream().cuda_stream)


def run():
    rc = L.pu_tc_linear_fwd(x.data_ptr(), a.K, wt.data_ptr(), a.K, None, y.data_ptr(), a.N, a.M, a.K, a.N, int(a.accumulate), None,
                            None, a.mode, flag.data_ptr(), ws.data_ptr(), ws.numel(), st)
    assert rc == 0, rc


for _ in range(3):
    run()
torch.cuda.synchronize()
L.pu_tc_debug_set_timeline(tl.data_ptr())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run(); e1.record()
torch.cuda.synchronize()
L.pu_tc_debug_set_timeline(None)
assert int(flag.item()) == 0, "barrier wait timed out"
t = tl.view(R, I, E).cpu()
t0 = int(t[t > 0].min())
names = ["loader", "converter", "issuer", "epilogue0", "epilogue1"]
print(f"M={a.M} K={a.K} N={a.N} accumulate={a.accumulate}: {e0.elapsed_time(e1) * 1e3:.1f} us; clock64 relative to the first stamp of CTA (0,0)")
for ri, name in enumerate(names):
    rows = [(i, [int(v) - t0 if v > 0 else -1 for v in t[ri, i]]) for i in range(min(I, a.items)) if int(t[ri, i].max()) > 0]
    if not rows:
        continue
    print(f"-- {name}: item, e0..e3, (e0 - previous e0)")
    prev = None
    for i, vals in rows:
        d = vals[0] - prev if prev is not None and vals[0] >= 0 else 0
        prev = vals[0] if vals[0] >= 0 else prev
        print(f"   {i:4d}  " + "  ".join(f"{v:9d}" for v in vals) + f"   +{d}")
    firsts = [vals[0] for _, vals in rows if vals[0] >= 0]
    if len(firsts) > 2:
        print(f"   average cycles per item: {(firsts[-1] - firsts[0]) / (len(firsts) - 1):.0f}")
