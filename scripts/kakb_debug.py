"""Pipeline cycle accounting of the layer-2 backward kernels (tc_ka.cuh, tc_kb.cuh) of the LAST tower of a train step
(the T-Net tower).  Needs a PGPD_DEBUG build of the library: PGPD_LIB=build/variants/libpgpd_dbg.so."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pointnetgpd_b200 import _abi as A
if os.environ.get("PGPD_LIB"):
    A.LIB_PATH = os.path.abspath(os.environ["PGPD_LIB"])
from pointnetgpd_b200 import synth as W
from pointnetgpd_b200.model.pointnet import PointNetCls
B, N = 512, 1024
m = PointNetCls(N, 3, 2); m.load_state_dict({k: torch.tensor(v) for k, v in W.make_state(0, k=2).items()}); m = m.cuda().train()
x = torch.tensor(W.make_clouds(1, B, N, "box")).cuda()
y = torch.tensor(W.make_labels(2, B, 2)).cuda()
lib = A.load()
def step():
    m.zero_grad()
    logp, _ = m(x)
    torch.nn.functional.nll_loss(logp, y).backward()
for _ in range(2):
    step()
torch.cuda.synchronize()
lib.pgpd_debug_stream_counters(1)
step()
torch.cuda.synchronize()
buf = (ctypes.c_longlong * (512 * 8))()
lib.pgpd_debug_l3_counters.argtypes = [ctypes.POINTER(ctypes.c_longlong)]
lib.pgpd_debug_l3_counters(buf)
lib.pgpd_debug_stream_counters(0)
a = np.array(buf[:], dtype=np.int64).reshape(512, 8)
names = ["loader wait buf_empty", "load latency", "converter work", "mma wait op_ready", "mma issue", "epi wait acc_full", "epi work", "total"]
tiles = B * (N // 64) / 148
for kname, rows in (("k_ka_tc", a[:148]), ("k_kb_tc", a[256:256 + 148])):
    print("kernel:", kname)
    for i, n in enumerate(names):
        print("  %-24s mean %10.0f cycles/CTA   %8.0f per tile" % (n, rows[:, i].mean(), rows[:, i].mean() / tiles))
