"""Pipeline cycle accounting of the streaming tcgen05 kernels: runs a train step with only ONE of them enabled
as tensor-core kernel at a time (PGPD_TC_MASK) would change numerics, so instead we run the full step and read the
counters of the LAST stream kernel launched; select it with the env var WHICH=(fwd|bwda|bwdb)."""
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
y = torch.tensor(W.make_labels(2, B, 2)).cuda()
lib = A.load()
which = os.environ.get("WHICH", "bwdb")
def step():
    m.zero_grad()
    logp, _ = m(x)
    torch.nn.functional.nll_loss(logp, y).backward()
for _ in range(2):
    step()
torch.cuda.synchronize()
lib.pgpd_debug_stream_counters(1)
if which == "fwd":
    with torch.no_grad():
        m(x)          # last stream kernel = layer-2 forward of the trunk
else:
    step()            # last stream kernel of a step = L2BwdB of the STN tower
torch.cuda.synchronize()
buf = (ctypes.c_longlong * (256 * 8))()
lib.pgpd_debug_l3_counters.argtypes = [ctypes.POINTER(ctypes.c_longlong)]
lib.pgpd_debug_l3_counters(buf)
lib.pgpd_debug_stream_counters(0)
a = np.array(buf[:], dtype=np.int64).reshape(256, 8)[:148]
names = ["mma wait b_full", "mma wait tmem_empty", "mma total", "prod wait", "prod work", "epi wait", "epi work"]
tiles = 4096 / 148
print("kernel:", which)
for i, n in enumerate(names):
    print("%-22s mean %10.0f cycles/CTA   %8.0f per tile" % (n, a[:, i].mean(), a[:, i].mean() / tiles))
