"""Dev tool: run one train step and dump the gradients (np.savez) -- used to A/B a single kernel via PGPD_TC_MASK."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import weights as W
from pointnetgpd_b200.model.pointnet import PointNetCls
out = sys.argv[1]
B, N, k = int(os.environ.get("B", 24)), int(os.environ.get("N", 1000)), 3
st = W.make_state(960, k=k, style="wild")
m = PointNetCls(N, 3, k); m.load_state_dict({kk: torch.tensor(v) for kk, v in st.items()}); m = m.cuda().train()
x = torch.tensor(W.make_clouds(961, B, N, "dup")).cuda()
y = torch.tensor(W.make_labels(962, B, k)).cuda()
logp, trans = m(x)
torch.nn.functional.nll_loss(logp, y).backward()
torch.cuda.synchronize()
np.savez(out, logp=logp.detach().cpu().numpy(), **{n: p.grad.cpu().numpy() for n, p in m.named_parameters()})
print("saved", out)
