"""Per-kernel device times of an eval forward (BASELINE config 5 shape by default)."""
import collections, os, re, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pointnetgpd_b200 import synth as W
from pointnetgpd_b200 import _abi as A
if os.environ.get("PGPD_LIB"):
    A.LIB_PATH = os.path.abspath(os.environ["PGPD_LIB"])
from pointnetgpd_b200.model.pointnet import PointNetCls
B, N, k = int(os.environ.get("B", 4096)), int(os.environ.get("N", 750)), 2
m = PointNetCls(N, 3, k); m.load_state_dict({kk: torch.tensor(v) for kk, v in W.make_state(0, k=k, style="wild").items()}); m = m.cuda().eval()
x = torch.tensor(W.make_clouds(5, B, N, "dup")).cuda()
with torch.no_grad():
    for _ in range(3): out = m(x)
    torch.cuda.synchronize()
    print("lib", A.LIB_PATH, "checksum %.6f %.6f" % (float(out[0].double().sum()), float(out[1].double().abs().sum())))
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3): m(x)
        torch.cuda.synchronize()
agg = collections.OrderedDict()
for ev in prof.events():
    if ev.device_type is not None and "cuda" in str(ev.device_type).lower():
        nm = re.sub(r"\(.*", "", ev.name)[:70]
        a = agg.setdefault(nm, [0, 0.0]); a[0] += 1; a[1] += ev.device_time_total
tot = sum(v[1] for v in agg.values())
for nm, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
    print("%9.1f us/fwd %4d  %5.1f%%  %s" % (t / 3, n // 3, 100 * t / tot, nm))
print("sum of kernels %.1f us/fwd (B=%d N=%d)" % (tot / 3, B, N))
