"""GPU test: the CUDA-graph captured training step reproduces the eager step exactly."""
import pytest
import torch
import torch.nn.functional as F

from oracle import weights as W
from pointnetgpd_b200.graph import GraphedTrainStep
from pointnetgpd_b200.model.pointnet import PointNetCls

pytestmark = pytest.mark.gpu


def _make(N, k, seed=970):
    st = W.make_state(seed, k=k)
    m = PointNetCls(num_points=N, k=k)
    m.load_state_dict({kk: torch.tensor(v) for kk, v in st.items()})
    return m.cuda().train()


def test_graphed_step_matches_eager_steps():
    B, N, k = 32, 300, 2
    xs = [torch.tensor(W.make_clouds(971 + i, B, N, "box")).cuda() for i in range(4)]
    ys = [torch.tensor(W.make_labels(981 + i, B, k)).cuda() for i in range(4)]
    # eager reference: 4 steps from the initial state (GraphedTrainStep's warm-up steps are rolled back: parameters, BatchNorm
    # buffers and optimizer state are restored before capture, so constructing it does not change the training trajectory)
    m0 = _make(N, k)
    o0 = torch.optim.Adam(m0.parameters(), lr=0.005, fused=True, capturable=True)
    def eager(m, o, x, y):
        o.zero_grad(set_to_none=True)
        logp, _ = m(x)
        loss = F.nll_loss(logp, y)
        loss.backward()
        o.step()
        return loss.detach().clone()
    ref = [eager(m0, o0, x, y) for x, y in zip(xs, ys)]
    m1 = _make(N, k)
    o1 = torch.optim.Adam(m1.parameters(), lr=0.005, fused=True, capturable=True)
    g = GraphedTrainStep(m1, o1, xs[0], ys[0], warmup=3)
    got = [g.step(x, y).clone() for x, y in zip(xs, ys)]
    torch.cuda.synchronize()
    for a, b in zip(ref, got):
        assert torch.equal(a, b)
    for p0, p1 in zip(m0.parameters(), m1.parameters()):
        assert torch.equal(p0, p1)
    assert int(m1.bn1.num_batches_tracked) == int(m0.bn1.num_batches_tracked) == 4


def test_staged_input_steps_match_plain_steps():
    """The prefetching input path (staging.StagedInput: next batch's H2D copy on a side stream while a step runs) feeds the
    captured step the same batches as the plain path."""
    B, N, k = 16, 200, 2
    xs = [torch.tensor(W.make_clouds(991 + i, B, N, "box")).pin_memory() for i in range(5)]
    ys = [torch.tensor(W.make_labels(995 + i, B, k)).pin_memory() for i in range(5)]
    losses = []
    for staged in (False, True):
        m = _make(N, k)
        o = torch.optim.Adam(m.parameters(), lr=0.005, fused=True, capturable=True)
        g = GraphedTrainStep(m, o, xs[0].cuda(), ys[0].cuda(), warmup=3)
        out = []
        if staged:
            st = g.staged_input()
            st.prefetch((xs[0], ys[0]))
            for i in range(5):
                loss = g.step_staged()
                if i + 1 < 5:
                    st.prefetch((xs[i + 1], ys[i + 1]))
                out.append(float(loss.item()))
        else:
            for x, y in zip(xs, ys):
                out.append(float(g.step(x, y).item()))
        losses.append(out)
    assert losses[0] == losses[1]
    assert len(set(losses[0])) > 1
