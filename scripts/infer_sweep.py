"""Inference throughput sweep (BASELINE config 5: batched candidate scoring, 750 points, eval mode) + the deploy shape
B=1 (kinect2grasp.py:479), with the eager-PyTorch reference module of the same architecture (oracle torch port on the
GPU, TF32 off and on) beside it."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import weights as W, pointnet_torch_port as PT
from pointnetgpd_b200.model.pointnet import PointNetCls

def timeit(fn, n):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

N, k = 750, 2
st = W.make_state(0, k=k, style="wild")
m = PointNetCls(N, 3, k); m.load_state_dict({kk: torch.tensor(v) for kk, v in st.items()}); m = m.cuda().eval()
sd = {kk: v.cuda() for kk, v in PT.to_torch_state(st, torch.float32).items()}
rows = []
for B in (1, 64, 512, 4096):
    x = torch.tensor(W.make_clouds(5, B, N, "dup")).cuda()
    with torch.no_grad():
        ours = timeit(lambda: m(x), 30 if B < 4096 else 10)
        g = torch.cuda.CUDAGraph()
        m(x)
        with torch.cuda.graph(g):
            out = m(x)
        ours_graph = timeit(g.replay, 30 if B < 4096 else 10)
        res = {}
        for tf32 in (False, True):
            torch.backends.cudnn.allow_tf32 = tf32; torch.backends.cuda.matmul.allow_tf32 = tf32
            res[tf32] = timeit(lambda: PT.pointnetcls_forward(sd, x, training=False), 10 if B < 4096 else 3)
    rows.append({"B": B, "N": N, "ours_ms": ours, "ours_graph_ms": ours_graph, "ours_grasps_per_s": B / ours_graph * 1e3,
                 "eager_torch_fp32_ms": res[False], "eager_torch_tf32_ms": res[True],
                 "eager_torch_fp32_grasps_per_s": B / res[False] * 1e3})
    print(json.dumps(rows[-1]))
# training step, eager PyTorch on the same GPU (the "library Blackwell" bar of BASELINE.md section 3)
B, N = 512, 1024
st = W.make_state(0, k=2)
x = torch.tensor(W.make_clouds(1, B, N, "box")).cuda(); y = torch.tensor(W.make_labels(2, B, 2)).cuda()
for tf32 in (False, True):
    torch.backends.cudnn.allow_tf32 = tf32; torch.backends.cuda.matmul.allow_tf32 = tf32
    sdt = {kk: v.cuda() for kk, v in PT.to_torch_state(st, torch.float32).items()}
    for kk, v in sdt.items():
        if v.is_floating_point() and not kk.endswith(("running_mean", "running_var")): v.requires_grad_(True)
    params = [v for v in sdt.values() if v.requires_grad]
    opt = torch.optim.Adam(params, lr=0.005, fused=True)
    def step():
        opt.zero_grad(set_to_none=True)
        logp, _ = PT.pointnetcls_forward(sdt, x, training=True)
        torch.nn.functional.nll_loss(logp, y).backward()
        opt.step()
    ms = timeit(step, 5)
    print(json.dumps({"train_step_eager_torch": True, "tf32": tf32, "B": B, "N": N, "ms_per_step": ms, "grasps_per_s": B / ms * 1e3}))
