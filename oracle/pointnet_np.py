"""numpy restatement (any dtype, normally float64) of the reference PointNetCls
forward AND backward, with every derivative written out by hand.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  The product never imports this.

Purpose: an autograd-free fp64 ground truth for error budgeting (SURVEY.md
section 7.2 C: gradient parity is judged against fp64, because arg-max routing makes
fp32 gradients discontinuous), and an independent check of the torch port.
Dense and simple on purpose -- O(B*N*1024) memory -- so only for small cases.

Reference lines (relative to /root/reference/PointNetGPD/):
  STN3d.forward         model/pointnet.py:27-45
  PointNetfeat.forward  model/pointnet.py:137-151
  PointNetCls.forward   model/pointnet.py:189-194
  loss                  main_1v.py:74  (F.nll_loss, mean reduction)
Maths: SURVEY.md Appendix A.
"""
import numpy as np

EPS = 1e-5
MOMENTUM = 0.1


# ----------------------------------------------------------------------------- building blocks
def _bn_fwd(y, g, be, rm, rv, training, axis):
    """BatchNorm1d over `axis` (tuple of reduced axes); y has channels on axis 1."""
    shp = [1] * y.ndim
    shp[1] = -1
    if training:
        n = y.size // y.shape[1]
        if n <= 1:
            raise ValueError("Expected more than 1 value per channel when training")
        mu = y.mean(axis=axis)
        var = ((y - mu.reshape(shp)) ** 2).mean(axis=axis)            # biased
        new_rm = (1 - MOMENTUM) * rm + MOMENTUM * mu
        new_rv = (1 - MOMENTUM) * rv + MOMENTUM * var * n / (n - 1)   # unbiased for the running stat
    else:
        mu, var, new_rm, new_rv = rm, rv, rm, rv
    r = 1.0 / np.sqrt(var + EPS)
    yhat = (y - mu.reshape(shp)) * r.reshape(shp)
    z = g.reshape(shp) * yhat + be.reshape(shp)
    return z, (yhat, r), (new_rm, new_rv)


def _bn_bwd(dz, g, cache, training, axis):
    yhat, r = cache
    shp = [1] * dz.ndim
    shp[1] = -1
    dg = (dz * yhat).sum(axis=axis)
    dbe = dz.sum(axis=axis)
    if training:
        n = dz.size // dz.shape[1]
        dy = (g * r).reshape(shp) * (dz - (dbe / n).reshape(shp) - yhat * (dg / n).reshape(shp))
    else:
        dy = (g * r).reshape(shp) * dz
    return dy, dg, dbe


def _conv_fwd(a, W, b):
    """Conv1d(k=1): a [B,Cin,N], W [Cout,Cin(,1)] -> [B,Cout,N]  (pointnet.py:12-14,127-129)."""
    W2 = W.reshape(W.shape[0], -1)
    return np.einsum("oc,bcn->bon", W2, a) + b.reshape(1, -1, 1)


def _conv_bwd(dy, a, W):
    W2 = W.reshape(W.shape[0], -1)
    dW = np.einsum("bon,bcn->oc", dy, a).reshape(W.shape)
    db = dy.sum(axis=(0, 2))
    da = np.einsum("oc,bon->bcn", W2, dy)
    return da, dW, db


def _lin_fwd(a, W, b):
    return a @ W.T + b


def _lin_bwd(dy, a, W):
    return dy @ W, dy.T @ a, dy.sum(axis=0)


def _maxpool_fwd(o):
    """MaxPool1d(N) over the last axis; FIRST arg-max on ties (pointnet.py:32,148)."""
    idx = o.argmax(axis=2)          # numpy argmax returns the first maximal index
    g = np.take_along_axis(o, idx[:, :, None], axis=2)[:, :, 0]
    return g, idx


def _maxpool_bwd(dg, idx, N):
    do = np.zeros(dg.shape + (N,), dtype=dg.dtype)
    np.put_along_axis(do, idx[:, :, None], dg[:, :, None], axis=2)
    return do


# ----------------------------------------------------------------------------- towers / heads
def _tower_fwd(sd, p, x, training, relu_last, new_stats):
    """3->64->128->1024 Conv1d+BN(+ReLU) tower and max-pool.
    pointnet.py:29-33 (relu_last=True) / :144-149 (relu_last=False)."""
    cache = {}
    a = x
    for l, relu in ((1, True), (2, True), (3, relu_last)):
        c, bn = p + "conv%d" % l, p + "bn%d" % l
        y = _conv_fwd(a, sd[c + ".weight"], sd[c + ".bias"])
        z, bc, st = _bn_fwd(y, sd[bn + ".weight"], sd[bn + ".bias"], sd[bn + ".running_mean"],
                            sd[bn + ".running_var"], training, (0, 2))
        new_stats[bn] = st
        o = np.maximum(z, 0) if relu else z
        cache[l] = (a, bc, z, relu)
        a = o
    g, idx = _maxpool_fwd(a)
    cache["idx"], cache["N"] = idx, x.shape[2]
    return g, cache


def _tower_bwd(sd, p, dg, cache, training, grads):
    do = _maxpool_bwd(dg, cache["idx"], cache["N"])
    for l in (3, 2, 1):
        a, bc, z, relu = cache[l]
        c, bn = p + "conv%d" % l, p + "bn%d" % l
        dz = do * (z > 0) if relu else do
        dy, dgam, dbe = _bn_bwd(dz, sd[bn + ".weight"], bc, training, (0, 2))
        grads[bn + ".weight"], grads[bn + ".bias"] = dgam, dbe
        do, dW, db = _conv_bwd(dy, a, sd[c + ".weight"])
        grads[c + ".weight"], grads[c + ".bias"] = dW, db
    return do  # gradient w.r.t. the tower input [B,3,N]


def _head_fwd(sd, p, g, bn_names, training, new_stats):
    """Linear+BN+ReLU x2 then Linear.  pointnet.py:35-37 (STN3d) / :191-193 (PointNetCls)."""
    cache = {}
    a = g
    for l in (1, 2):
        fc, bn = p + "fc%d" % l, p + bn_names[l - 1]
        y = _lin_fwd(a, sd[fc + ".weight"], sd[fc + ".bias"])
        z, bc, st = _bn_fwd(y, sd[bn + ".weight"], sd[bn + ".bias"], sd[bn + ".running_mean"],
                            sd[bn + ".running_var"], training, (0,))
        new_stats[bn] = st
        cache[l] = (a, bc, z)
        a = np.maximum(z, 0)
    cache[3] = a
    return _lin_fwd(a, sd[p + "fc3.weight"], sd[p + "fc3.bias"]), cache


def _head_bwd(sd, p, dout, cache, bn_names, training, grads):
    da, dW, db = _lin_bwd(dout, cache[3], sd[p + "fc3.weight"])
    grads[p + "fc3.weight"], grads[p + "fc3.bias"] = dW, db
    for l in (2, 1):
        a, bc, z = cache[l]
        fc, bn = p + "fc%d" % l, p + bn_names[l - 1]
        dz = da * (z > 0)
        dy, dgam, dbe = _bn_bwd(dz, sd[bn + ".weight"], bc, training, (0,))
        grads[bn + ".weight"], grads[bn + ".bias"] = dgam, dbe
        da, dW, db = _lin_bwd(dy, a, sd[fc + ".weight"])
        grads[fc + ".weight"], grads[fc + ".bias"] = dW, db
    return da


# ----------------------------------------------------------------------------- whole model
def forward(sd, x, training=False):
    """PointNetCls.forward.  Returns (logp, trans, cache, new_running_stats)."""
    new_stats = {}
    g_stn, c_stn_t = _tower_fwd(sd, "feat.stn.", x, training, True, new_stats)          # :29-33
    t9, c_stn_h = _head_fwd(sd, "feat.stn.", g_stn, ("bn4", "bn5"), training, new_stats)  # :35-37
    trans = (t9 + np.eye(3, dtype=x.dtype).reshape(1, 9)).reshape(-1, 3, 3)             # :39-44
    xt = np.einsum("bjn,bji->bin", x, trans)                                            # :140-143  x' = T^T x
    g, c_t = _tower_fwd(sd, "feat.", xt, training, False, new_stats)                    # :144-149
    logits, c_h = _head_fwd(sd, "", g, ("bn1", "bn2"), training, new_stats)             # :191-193
    m = logits.max(axis=1, keepdims=True)
    logp = logits - m - np.log(np.exp(logits - m).sum(axis=1, keepdims=True))           # :194
    cache = dict(stn_t=c_stn_t, stn_h=c_stn_h, t=c_t, h=c_h, x=x, logp=logp)
    return logp, trans, cache, new_stats


def backward(sd, cache, dlogp, dtrans_ext=None, training=True):
    """Gradients of sum(dlogp*logp) + sum(dtrans_ext*trans) w.r.t. every parameter."""
    grads = {}
    x, logp = cache["x"], cache["logp"]
    dlogits = dlogp - np.exp(logp) * dlogp.sum(axis=1, keepdims=True)
    dg = _head_bwd(sd, "", dlogits, cache["h"], ("bn1", "bn2"), training, grads)
    dxt = _tower_bwd(sd, "feat.", dg, cache["t"], training, grads)
    dT = np.einsum("bjn,bin->bji", x, dxt)
    if dtrans_ext is not None:
        dT = dT + dtrans_ext
    dg_stn = _head_bwd(sd, "feat.stn.", dT.reshape(-1, 9), cache["stn_h"], ("bn4", "bn5"), training, grads)
    _tower_bwd(sd, "feat.stn.", dg_stn, cache["stn_t"], training, grads)
    return grads


def nll_train_step(sd, x, target, training=True):
    """forward + mean NLL + backward (main_1v.py:72-75)."""
    logp, trans, cache, new_stats = forward(sd, x, training)
    B = x.shape[0]
    loss = -logp[np.arange(B), target].mean()
    dlogp = np.zeros_like(logp)
    dlogp[np.arange(B), target] = -1.0 / B
    grads = backward(sd, cache, dlogp, None, training)
    return logp, trans, loss, grads, new_stats


def cast_state(np_state, dtype):
    return {k: (v if k.endswith("num_batches_tracked") else np.asarray(v, dtype=dtype)) for k, v in np_state.items()}
