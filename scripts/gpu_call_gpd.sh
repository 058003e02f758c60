#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpd.py -m gpu -q --timeout=600 2>&1 | tail -4
python - <<'PY'
import time, torch, sys, collections, re
sys.path.insert(0, '.')
from pointnetgpd_b200.model.gpd import GPDClassifier
B=512
m = GPDClassifier(3).cuda().train()
x = torch.randn(B, 3, 60, 60, device='cuda'); y = torch.randint(0, 2, (B,), device='cuda')
def step():
    m.zero_grad(); l = torch.nn.functional.nll_loss(m(x), y); l.backward()
for _ in range(3): step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3): step()
    torch.cuda.synchronize()
agg = collections.OrderedDict()
for ev in prof.events():
    if ev.device_type is not None and "cuda" in str(ev.device_type).lower():
        nm = re.sub(r"\(.*", "", ev.name)[:60]
        a = agg.setdefault(nm, [0, 0.0]); a[0] += 1; a[1] += ev.device_time_total
for nm, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
    print("%9.1f us/step %3d  %s" % (t / 3, n // 3, nm))
PY
