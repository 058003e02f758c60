"""One small train step (forward + backward through libpgpd, default tcgen05 dispatch) for compute-sanitizer runs:
    compute-sanitizer --tool racecheck|synccheck|memcheck python scripts/sanitize_step.py [B N]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pointnetgpd_b200 import synth as W
from pointnetgpd_b200.model.pointnet import PointNetCls

B = int(sys.argv[1]) if len(sys.argv) > 1 else 6
N = int(sys.argv[2]) if len(sys.argv) > 2 else 700
st = W.make_state(7, k=2)
m = PointNetCls(num_points=N, k=2)
m.load_state_dict({k: torch.tensor(v) for k, v in st.items()})
m = m.cuda().train()
x = torch.tensor(W.make_clouds(8, B, N, "box")).cuda()
y = torch.tensor(W.make_labels(9, B, 2)).cuda()
logp, trans = m(x)
loss = torch.nn.functional.nll_loss(logp, y)
loss.backward()
torch.cuda.synchronize()
m.eval()
with torch.no_grad():
    lp2, _ = m(x)
torch.cuda.synchronize()
print("sanitize_step ok: loss %.6f, |grad fc3| %.4e, eval logp[0] %s" % (float(loss.detach()), float(m.fc3.weight.grad.norm()), lp2[0].tolist()))
