"""Per-kernel device times of an eager train step (torch.profiler / CUPTI), for A/B-ing kernel variants.
Dev tool: PGPD_LIB selects an alternative build of libpgpd (never used by the product path)."""
import collections, os, re, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import weights as W
from pointnetgpd_b200 import _abi as A
if os.environ.get("PGPD_LIB"):
    A.LIB_PATH = os.path.abspath(os.environ["PGPD_LIB"])
from pointnetgpd_b200.model.pointnet import PointNetCls
B, N, k = int(os.environ.get("B", 512)), int(os.environ.get("N", 1024)), 2
st = W.make_state(0, k=k)
m = PointNetCls(N, 3, k); m.load_state_dict({kk: torch.tensor(v) for kk, v in st.items()}); m = m.cuda().train()
x = torch.tensor(W.make_clouds(1, B, N, "box")).cuda()
y = torch.tensor(W.make_labels(2, B, k)).cuda()
def step():
    m.zero_grad(set_to_none=True)
    logp, _ = m(x)
    loss = torch.nn.functional.nll_loss(logp, y)
    loss.backward()
    return loss
for _ in range(3):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    l = step()
e1.record(); torch.cuda.synchronize()
print("lib", A.LIB_PATH, "eager fwd+bwd ms/step %.3f loss %.6f" % (e0.elapsed_time(e1) / 10, float(l)))
try:
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            step()
        torch.cuda.synchronize()
    agg = collections.OrderedDict()
    for ev in prof.events():
        if ev.device_type is not None and "cuda" in str(ev.device_type).lower():
            nm = re.sub(r"\(.*", "", ev.name)[:70]
            a = agg.setdefault(nm, [0, 0.0]); a[0] += 1; a[1] += ev.device_time_total if hasattr(ev, "device_time_total") else ev.cuda_time_total
    tot = sum(v[1] for v in agg.values())
    for nm, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(os.environ.get("TOP", 24))]:
        print("%9.1f us/step %4d  %5.1f%%  %s" % (t / 3, n // 3, 100 * t / tot, nm))
    print("sum of kernels %.1f us/step" % (tot / 3))
except Exception as e:
    print("profiler unavailable:", type(e).__name__, e)
