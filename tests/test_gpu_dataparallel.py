"""The reference's own multi-GPU mode: nn.DataParallel(model, device_ids=[...]) (main_1v.py:163-165) -- single process,
one Python thread per device calling forward concurrently, gradients reduced onto device 0.  Needs >= 2 GPUs."""
import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from oracle import pointnet_torch_port as PT
from oracle import weights as W
from pointnetgpd_b200.model.pointnet import PointNetCls

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_dataparallel_matches_per_shard_oracle():
    B, N, k = 64, 300, 2
    st = W.make_state(990, k=k)
    m = PointNetCls(num_points=N, k=k)
    m.load_state_dict({kk: torch.tensor(v) for kk, v in st.items()})
    dp = nn.DataParallel(m, device_ids=[0, 1]).cuda()
    dp.train()
    x = torch.tensor(W.make_clouds(991, B, N, "box")).cuda()
    y = torch.tensor(W.make_labels(992, B, k)).cuda()
    logp, trans = dp(x)
    assert logp.shape == (B, k) and trans.shape == (B, 3, 3) and logp.device.index == 0
    F.nll_loss(logp, y).backward()
    # DataParallel semantics: each replica normalises over its own shard (B/2 clouds), gradients are summed
    ref_logp, grads = [], None
    for sh in range(2):
        sd = PT.to_torch_state(st, torch.float64, requires_grad=True)
        xs = x[sh * B // 2:(sh + 1) * B // 2].double().cpu()
        lp, _ = PT.pointnetcls_forward(sd, xs, training=True)
        ref_logp.append(lp.detach())
        # loss = mean over the FULL batch -> each shard contributes sum/B
        (-lp[torch.arange(B // 2), y[sh * B // 2:(sh + 1) * B // 2].cpu()].sum() / B).backward()
        g = {kk: v.grad for kk, v in sd.items() if v.requires_grad}
        grads = g if grads is None else {kk: grads[kk] + g[kk] for kk in g}
    ref_logp = torch.cat(ref_logp)
    assert float((logp.detach().cpu().double() - ref_logp).abs().max()) < 1e-3
    for n, p in m.named_parameters():
        r = grads[n].reshape(p.shape)
        if float(r.norm()) < 1e-8:
            continue
        rel = float((p.grad.cpu().double() - r).norm() / r.norm())
        assert rel < 2e-2, (n, rel)
