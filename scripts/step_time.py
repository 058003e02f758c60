"""Graph-replayed train-step time (as bench.py measures it) for a given build of the library.
Dev tool: PGPD_LIB selects an alternative build of libpgpd (never used by the product path)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pointnetgpd_b200 import _abi as A
if os.environ.get("PGPD_LIB"):
    A.LIB_PATH = os.path.abspath(os.environ["PGPD_LIB"])
from pointnetgpd_b200 import synth as W
from pointnetgpd_b200.graph import GraphedTrainStep
from pointnetgpd_b200.model.pointnet import PointNetCls
B, N, k = int(os.environ.get("B", 512)), int(os.environ.get("N", 1024)), 2
m = PointNetCls(N, 3, k); m.load_state_dict({kk: torch.tensor(v) for kk, v in W.make_state(0, k=k).items()}); m = m.cuda().train()
opt = torch.optim.Adam(m.parameters(), lr=0.005, fused=True, capturable=True)
xs = [torch.tensor(W.make_clouds(1234 + i, B, N, "box")).cuda() for i in range(8)]
ys = [torch.tensor(W.make_labels(4321 + i, B, k)).cuda() for i in range(8)]
g = GraphedTrainStep(m, opt, xs[0], ys[0], warmup=3)
for i in range(5): g.step(xs[i % 8], ys[i % 8])
torch.cuda.synchronize()
best = 1e9
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(20): loss = g.step(xs[i % 8], ys[i % 8])
    e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / 20)
print("lib %s graph step %.4f ms (best of 3 x 20), loss %.6f" % (os.path.basename(A.LIB_PATH), best, float(loss)))
