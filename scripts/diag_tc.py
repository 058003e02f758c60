"""Per-parameter comparison of a train step through the CUDA-core kernels vs the tcgen05 dispatch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import weights as W
from pointnetgpd_b200 import _abi as A
from pointnetgpd_b200.functional import run_module
from pointnetgpd_b200.model.pointnet import PointNetCls
B, N, k = int(sys.argv[1]), int(sys.argv[2]), 3
st = W.make_state(960, k=k, style="wild")
x = torch.tensor(W.make_clouds(961, B, N, "dup")).cuda()
y = torch.tensor(W.make_labels(962, B, k)).cuda()
res = []
for extra in (A.F_SIMT, 0, 0):
    m = PointNetCls(N, 3, k); m.load_state_dict({kk: torch.tensor(v) for kk, v in st.items()}); m = m.cuda().train()
    logp, trans = run_module(m, A.PGPD_CLS, x, k=k, flags_extra=extra)
    torch.nn.functional.nll_loss(logp, y).backward()
    torch.cuda.synchronize()
    res.append((logp.detach(), {n: p.grad.clone() for n, p in m.named_parameters()}))
print("mask", os.environ.get("PGPD_TC_MASK"), "B", B, "N", N, "logp diff", float((res[0][0] - res[1][0]).abs().max()))
for n in res[0][1]:
    g0, g1, g2 = res[0][1][n], res[1][1][n], res[2][1][n]
    rel = float((g0 - g1).norm() / g0.norm().clamp_min(1e-30))
    det = bool(torch.equal(g1, g2))
    flag = "" if (rel < 3e-2 or float(g0.norm()) < 1e-5) and det else "   <<<<<<"
    if flag or "-v" in sys.argv:
        print("  %-26s |simt| %.3e rel %.3e deterministic=%s%s" % (n, float(g0.norm()), rel, det, flag))
