"""Does capturing the whole train step in a CUDA graph help?  (side experiment)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from oracle import weights as W
from pointnetgpd_b200.model.pointnet import PointNetCls
B, N, k = 512, 1024, 2
st = W.make_state(0, k=k)
m = PointNetCls(N, 3, k); m.load_state_dict({kk: torch.tensor(v) for kk, v in st.items()}); m = m.cuda().train()
opt = torch.optim.Adam(m.parameters(), lr=0.005, fused=True, capturable=True)
xs = torch.tensor(W.make_clouds(1, B, N, "box")).cuda(); ys = torch.tensor(W.make_labels(2, B, k)).cuda()
sx, sy = xs.clone(), ys.clone()
def step():
    opt.zero_grad(set_to_none=True)
    logp, _ = m(sx)
    loss = F.nll_loss(logp, sy)
    loss.backward()
    opt.step()
    return loss
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(3):
        step()
torch.cuda.current_stream().wait_stream(s)
torch.cuda.synchronize()
def timeit(fn, n=20):
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
print("eager ms/step", timeit(step))
g = torch.cuda.CUDAGraph()
opt.zero_grad(set_to_none=True)
with torch.cuda.graph(g):
    loss = step()
print("graph ms/step", timeit(g.replay), "loss", float(loss))
