"""CPU tests of the CUDA-core kernels and the C++ host orchestration of libpgpd, run through the
SIMT emulator (tests/simt_emu), against the golden vectors and the fp64 numpy oracle.
No GPU needed.  (The same checks run on the real kernels in tests/test_gpu_parity.py.)"""
import ctypes as C

import numpy as np
import pytest

import emu_util as E
from golden_util import SMALL_CASES, load_case, grad_errors, is_zero_grad_param, param_names
from oracle import pointnet_np as PN
from oracle import weights as W
from pointnetgpd_b200 import _abi as A

LOGP_TOL = 1e-3      # north_star: outputs within 1e-3 fp32 of the reference
GRAD_FLOOR = 5e-3    # relative; arg-max routing makes fp32 gradients discontinuous


def _copy(st):
    return {k: v.copy() for k, v in st.items()}


@pytest.mark.parametrize("name", SMALL_CASES)
def test_eval_forward_matches_golden(name):
    c = load_case(name)
    r = E.run_model(_copy(c["state"]), c["x"], train=False)
    g = c["g"]
    assert np.abs(r["out"] - g["eval_logp_f64"]).max() < LOGP_TOL
    assert np.abs(r["trans"] - g["eval_trans_f64"]).max() < LOGP_TOL
    assert (r["out"].argmax(1) == g["eval_logp_f64"].argmax(1)).all()
    # eval mode must not touch the running statistics
    for k, v in c["state"].items():
        assert np.array_equal(np.asarray(v).reshape(-1), r["state"][k].reshape(-1)), k


@pytest.mark.parametrize("name", SMALL_CASES)
def test_train_step_matches_golden(name):
    c = load_case(name)
    g = c["g"]
    B, k = c["B"], c["k"]
    dlogp = np.zeros((B, k), np.float32)
    dlogp[np.arange(B), c["y"]] = -1.0 / B          # d(mean NLL)/d logp, main_1v.py:74
    r = E.run_model(_copy(c["state"]), c["x"], train=True, backward=True, dout=dlogp)
    assert np.abs(r["out"] - g["train_logp_f64"]).max() < LOGP_TOL
    assert np.abs(r["trans"] - g["train_trans_f64"]).max() < LOGP_TOL
    assert (r["out"].argmax(1) == g["train_logp_f64"].argmax(1)).all()
    ours = grad_errors(r["grads"], g, "f64")
    # the reference's own fp32 error w.r.t. fp64, from the golden file
    ref32 = {n: g[f"gsub_f32/{n}"] for n in r["grads"]}
    for n, (esub, enorm) in ours.items():
        if is_zero_grad_param(n):
            scale = max(float(g[f"gnorm_f64/fc3.weight"]), 1.0)
            assert np.abs(r["grads"][n]).max() < 1e-3 * scale, n
            continue
        ref_sub64 = g[f"gsub_f64/{n}"].astype(np.float64)
        ref_err = np.linalg.norm(ref32[n].astype(np.float64) - ref_sub64) / max(np.linalg.norm(ref_sub64), 1e-30)
        assert esub < max(GRAD_FLOOR, 10 * ref_err), (n, esub, ref_err)
        assert enorm < max(GRAD_FLOOR, 10 * ref_err), (n, enorm, ref_err)
    for key in A.buffer_keys():
        got = r["state"][key].reshape(-1)
        ref = np.asarray(g["buf_f64/" + key]).reshape(-1)
        if key.endswith("num_batches_tracked"):
            assert int(got[0]) == int(ref[0]) == 1
        else:
            assert np.abs(got - ref).max() < 1e-4 * max(1.0, np.abs(ref).max()), key


@pytest.mark.parametrize("name", SMALL_CASES[:1])
def test_general_output_gradients(name):
    """d(sum wl*logp + sum wt*trans): exercises the d(trans) input of pgpd_backward."""
    c = load_case(name)
    g = c["g"]
    r = E.run_model(_copy(c["state"]), c["x"], train=True, backward=True, dout=c["wl"], dtrans=c["wt"])
    errs = grad_errors(r["grads"], g, "f64", prefix="g2")
    for n, (esub, enorm) in errs.items():
        if is_zero_grad_param(n):
            continue
        ref32 = g[f"g2sub_f32/{n}"].astype(np.float64)
        ref64 = g[f"g2sub_f64/{n}"].astype(np.float64)
        ref_err = np.linalg.norm(ref32 - ref64) / max(np.linalg.norm(ref64), 1e-30)
        assert esub < max(GRAD_FLOOR, 10 * ref_err), (n, esub, ref_err)


@pytest.mark.parametrize("what", [A.PGPD_STN, A.PGPD_FEAT])
def test_stn_and_feat_modules(what):
    """STN3d and PointNetfeat as stand-alone modules (pointnet.py:27-45 / :137-151)."""
    B, N, k = 5, 67, 2
    st = W.make_state(21, k=k, style="wild")
    x = W.make_clouds(22, B, N, "box")
    sd64 = PN.cast_state(st, np.float64)
    ns = {}
    x64 = x.astype(np.float64)
    g_stn, c_t = PN._tower_fwd(sd64, "feat.stn.", x64, True, True, ns)
    t9, c_h = PN._head_fwd(sd64, "feat.stn.", g_stn, ("bn4", "bn5"), True, ns)
    trans = (t9 + np.eye(3).reshape(1, 9)).reshape(-1, 3, 3)
    wt = W.normal(23, (B, 3, 3))
    wg = W.normal(24, (B, 1024))
    grads = {}
    if what == A.PGPD_STN:
        dg = PN._head_bwd(sd64, "feat.stn.", wt.reshape(-1, 9), c_h, ("bn4", "bn5"), True, grads)
        PN._tower_bwd(sd64, "feat.stn.", dg, c_t, True, grads)
        r = E.run_model(_copy(st), x, what=what, train=True, backward=True, dtrans=wt)
    else:
        xt = np.einsum("bjn,bji->bin", x64, trans)
        G, c_tr = PN._tower_fwd(sd64, "feat.", xt, True, False, ns)
        dxt = PN._tower_bwd(sd64, "feat.", wg, c_tr, True, grads)
        dT = np.einsum("bjn,bin->bji", x64, dxt) + wt
        dg = PN._head_bwd(sd64, "feat.stn.", dT.reshape(-1, 9), c_h, ("bn4", "bn5"), True, grads)
        PN._tower_bwd(sd64, "feat.stn.", dg, c_t, True, grads)
        r = E.run_model(_copy(st), x, what=what, train=True, backward=True, dout=wg, dtrans=wt)
        assert np.abs(r["out"] - G).max() < 1e-3
    assert np.abs(r["trans"] - trans).max() < 1e-3
    for n, gv in r["grads"].items():
        ref = grads[n].reshape(gv.shape)
        nrm = np.linalg.norm(ref)
        if is_zero_grad_param(n) or nrm < 1e-9:
            continue
        assert np.linalg.norm(gv - ref) / nrm < 2e-2, n


@pytest.mark.parametrize("B,N", [(2, 1), (2, 129), (3, 257), (1, 40)])
def test_odd_shapes_eval(B, N):
    """tile tails (N not a multiple of 128), N=1, and B=1 in eval mode (kinect2grasp.py:479)."""
    st = W.make_state(31, k=3, style="wild")
    x = W.make_clouds(32, B, N, "dup" if N > 3 else "box")
    logp, trans, _, _ = PN.forward(PN.cast_state(st, np.float64), x.astype(np.float64), training=False)
    r = E.run_model(_copy(st), x, train=False)
    assert np.abs(r["out"] - logp).max() < LOGP_TOL
    assert np.abs(r["trans"] - trans).max() < LOGP_TOL


@pytest.mark.parametrize("B,N,kind", [(4, 150, "dup"), (3, 2, "box"), (8, 4, "box")])
def test_train_tail_tile_and_duplicates(B, N, kind):
    """Tile tails, duplicated points and -- N = 2, 4 -- clouds whose few points each own hundreds of arg-max channels: the rows of
    the sparse part of d a2 then span many warps of k_da2_sparse (head partials, whole-warp segments inside one row)."""
    k = 2
    st = W.make_state(41, k=k, style="wild")
    x = W.make_clouds(42, B, N, kind)
    sd64 = PN.cast_state(st, np.float64)
    logp, trans, cache, _ = PN.forward(sd64, x.astype(np.float64), training=True)
    wl = W.normal(43, (B, k))
    grads = PN.backward(sd64, cache, wl, None, True)
    r = E.run_model(_copy(st), x, train=True, backward=True, dout=wl)
    assert np.abs(r["out"] - logp).max() < LOGP_TOL
    for n, gv in r["grads"].items():
        ref = grads[n].reshape(gv.shape)
        nrm = np.linalg.norm(ref)
        if is_zero_grad_param(n) or nrm < 1e-9:
            continue
        assert np.linalg.norm(gv - ref) / nrm < 2e-2, n


def test_batch_of_one_in_train_mode_raises_like_batchnorm():
    st = W.make_state(51, k=2)
    x = W.make_clouds(52, 1, 16, "box")
    with pytest.raises(ValueError, match="more than 1 value per channel"):
        E.run_model(_copy(st), x, train=True)


def test_abi_argument_checks():
    lib = E.emu_lib()
    assert lib.pgpd_version() == 100
    assert lib.pgpd_workspace_bytes(A.PGPD_CLS, 0, 10, 2, 0) == 0
    st = W.make_state(61, k=2)
    stc = {k: np.ascontiguousarray(v).reshape(-1) if np.ndim(v) == 0 else np.ascontiguousarray(v) for k, v in st.items()}
    m = A.build_model(lambda key: stc[key].ctypes.data)
    x = W.make_clouds(62, 2, 8, "box")
    out = np.zeros((2, 2), np.float32)
    tr = np.zeros((2, 3, 3), np.float32)
    need = lib.pgpd_workspace_bytes(A.PGPD_CLS, 2, 8, 2, 0)
    ws = E.Guarded(need)
    # workspace too small
    rc = lib.pgpd_forward(A.PGPD_CLS, C.byref(m), x.ctypes.data, 2, 8, 2, 0, out.ctypes.data, tr.ctypes.data, ws.addr, need - 1, None)
    assert rc == A.E_WORKSPACE and b"small" in lib.pgpd_last_error()
    # misaligned workspace
    rc = lib.pgpd_forward(A.PGPD_CLS, C.byref(m), x.ctypes.data, 2, 8, 2, 0, out.ctypes.data, tr.ctypes.data, ws.addr + 4, need, None)
    assert rc == A.E_WORKSPACE
    # null input
    rc = lib.pgpd_forward(A.PGPD_CLS, C.byref(m), None, 2, 8, 2, 0, out.ctypes.data, tr.ctypes.data, ws.addr, need, None)
    assert rc == A.E_ARG
    # bad module id
    rc = lib.pgpd_forward(7, C.byref(m), x.ctypes.data, 2, 8, 2, 0, out.ctypes.data, tr.ctypes.data, ws.addr, need, None)
    assert rc == A.E_ARG
    # a conv weight that is only 4-byte aligned (the kernels read weights with 16-byte vector loads)
    big = np.zeros(stc["feat.conv3.weight"].size + 4, np.float32)
    off = 1 if big[1:].ctypes.data % 16 else 2
    mis = big[off:off + stc["feat.conv3.weight"].size]
    mis[:] = stc["feat.conv3.weight"].reshape(-1)
    assert mis.ctypes.data % 16 != 0
    table = dict(stc)
    table["feat.conv3.weight"] = mis
    m2 = A.build_model(lambda key: table[key].ctypes.data)
    rc = lib.pgpd_forward(A.PGPD_CLS, C.byref(m2), x.ctypes.data, 2, 8, 2, 0, out.ctypes.data, tr.ctypes.data, ws.addr, need, None)
    assert rc == A.E_ARG and b"aligned" in lib.pgpd_last_error()
    ws.check()


def test_permutation_and_duplicate_invariance_eval():
    """max-pool properties of the path: permuting points / appending duplicates leaves eval outputs unchanged
    (up to fp32 summation order in nothing -- the eval path has no cross-point reductions)."""
    st = W.make_state(71, k=2, style="wild")
    x = W.make_clouds(72, 3, 64, "box")
    r0 = E.run_model(_copy(st), x, train=False)
    perm = np.argsort(W.uniform01(73, 64))
    r1 = E.run_model(_copy(st), x[:, :, perm], train=False)
    assert np.array_equal(r0["out"], r1["out"])
    xd = np.concatenate([x, x[:, :, :17]], axis=2)
    r2 = E.run_model(_copy(st), xd, train=False)
    assert np.array_equal(r0["out"], r2["out"])


def test_two_phase_backward_equals_one_call():
    """pgpd_backward(PGPD_F_BWD_HEAD) followed by pgpd_backward(PGPD_F_BWD_STN) -- the split a data-parallel trainer uses to
    overlap the gradient exchange of the first half -- writes exactly the gradients of the single call."""
    c = load_case(SMALL_CASES[0])
    B, k = c["B"], c["k"]
    dlogp = np.zeros((B, k), np.float32)
    dlogp[np.arange(B), c["y"]] = -1.0 / B
    one = E.run_model(_copy(c["state"]), c["x"], train=True, backward=True, dout=dlogp, dtrans=c["wt"])
    two = E.run_model(_copy(c["state"]), c["x"], train=True, backward=True, dout=dlogp, dtrans=c["wt"], split_backward=True)
    for n in one["grads"]:
        assert np.array_equal(one["grads"][n], two["grads"][n]), n
