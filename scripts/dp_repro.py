"""1-GPU reproduction of the numerics of tests/test_gpu_dataparallel.py: each shard through the module, vs the fp64 oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from oracle import pointnet_torch_port as PT
from oracle import weights as W
from pointnetgpd_b200.model.pointnet import PointNetCls
B, N, k = 64, 300, 2
st = W.make_state(990, k=k)
x = torch.tensor(W.make_clouds(991, B, N, "box")).cuda()
y = torch.tensor(W.make_labels(992, B, k)).cuda()
for sh in range(2):
    m = PointNetCls(num_points=N, k=k); m.load_state_dict({kk: torch.tensor(v) for kk, v in st.items()}); m = m.cuda().train()
    xs = x[sh * B // 2:(sh + 1) * B // 2].contiguous(); ys = y[sh * B // 2:(sh + 1) * B // 2]
    lp, _ = m(xs)
    (-lp[torch.arange(B // 2), ys].sum() / B).backward()
    sd = PT.to_torch_state(st, torch.float64, requires_grad=True)
    rl, _ = PT.pointnetcls_forward(sd, xs.double().cpu(), training=True)
    (-rl[torch.arange(B // 2), ys.cpu()].sum() / B).backward()
    print("shard", sh, "max|dlogp|", float((lp.detach().cpu().double() - rl.detach()).abs().max()))
    for n, p in m.named_parameters():
        r = sd[n].grad.reshape(p.shape)
        if float(r.norm()) < 1e-8: continue
        rel = float((p.grad.cpu().double() - r).norm() / r.norm())
        if rel > 2e-3: print("   %-28s rel %.3e |ref| %.3e" % (n, rel, float(r.norm())))
