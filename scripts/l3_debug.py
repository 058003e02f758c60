"""Print the pipeline cycle accounting of the tcgen05 layer-3 kernel (run with PGPD_L3_DEBUG=1)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import weights as W
from pointnetgpd_b200 import _abi as A
if os.environ.get("PGPD_LIB"):
    A.LIB_PATH = os.path.abspath(os.environ["PGPD_LIB"])
from pointnetgpd_b200.model.pointnet import PointNetCls
B, N = 512, 1024
st = W.make_state(0, k=2)
m = PointNetCls(N, 3, 2); m.load_state_dict({k: torch.tensor(v) for k, v in st.items()}); m = m.cuda().train()
x = torch.tensor(W.make_clouds(1, B, N, "box")).cuda()
for _ in range(3):
    with torch.no_grad():
        m(x)
torch.cuda.synchronize()
lib = A.load()
print("lib", A.LIB_PATH)
buf = (ctypes.c_longlong * (512 * 8))()
lib.pgpd_debug_l3_counters.argtypes = [ctypes.POINTER(ctypes.c_longlong)]
lib.pgpd_debug_l3_counters(buf)
full = np.array(buf[:], dtype=np.int64).reshape(512, 8)
a = full[:148]
ver = "3"
if ver == "3":
    a = a[0::2]          # leader CTAs hold the MMA-loop counters
g = full[200:232]
g = g[g[:, 0] > 0]
print("version", ver, ": CTA wall time %.1f us, %.0f SM cycles -> effective SM clock %.3f GHz" % (g[:, 0].mean() / 1e3, g[:, 1].mean(), (g[:, 1] / g[:, 0]).mean()))
names = ["mma wait a2_full", "mma wait tmem_empty", "mma wait w_full (own)", "mma total", "mma wait w_full (peer relay)", "sum fetch latency of stalled stages", "stalled stages", "max fetch latency"]
tiles = 2048 / (74 if ver == "3" else 148)
for i, n in enumerate(names):
    print("%-22s mean %10.0f cycles/CTA   %8.0f per tile" % (n, a[:, i].mean(), a[:, i].mean() / tiles))

if a[:, 6].mean() > 0:
    print("mean fetch latency of a stalled stage: %.0f cycles (stalled %.1f of %d stages per tile)" % (a[:, 5].sum() / a[:, 6].sum(), a[:, 6].mean() / tiles, 8))
