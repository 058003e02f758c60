"""CPU port of the reference PointNetCls path, written against torch's functional ops.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  The product never imports this.

Why torch and not numpy: every arithmetic step of the reference path is a call
into PyTorch -- the reference's own CPU implementation IS
`torch.nn` on ATen/MKLDNN.  This port restates the same op sequence functionally
(no nn.Module, state passed as a dict), so that
  * it can be timed as the CPU arm (`bench.py --impl reference`, cpu_baseline
    kind "port") on the GPU box, where /root/reference does not exist, and
  * autograd through it gives reference-equivalent gradients (fp32 or fp64).
It is pinned against the imported reference module by oracle/make_golden.py
(tests/golden/*.npz) and tests/test_oracle.py.

Each function cites the reference lines it follows
(paths relative to /root/reference/PointNetGPD/).
"""
import torch
import torch.nn.functional as F

EPS = 1e-5        # nn.BatchNorm1d default eps      (model/pointnet.py:21-25,130-132,185-186)
MOMENTUM = 0.1    # nn.BatchNorm1d default momentum


def _bn(x, sd, name, training):
    """nn.BatchNorm1d forward incl. running-stat update (in place on sd).
    model/pointnet.py:29-31,35-36 (STN3d), :144-147 (PointNetfeat), :191-192 (PointNetCls)."""
    rm, rv = sd[name + ".running_mean"], sd[name + ".running_var"]
    if training:
        nbt = name + ".num_batches_tracked"
        if nbt in sd:
            sd[nbt] += 1
        n = x.numel() // x.shape[1]
        if n <= 1:
            # torch.nn.functional.batch_norm -> _verify_batch_size
            raise ValueError("Expected more than 1 value per channel when training, got input size {}".format(list(x.shape)))
    return F.batch_norm(x, rm, rv, sd[name + ".weight"], sd[name + ".bias"], training, MOMENTUM, EPS)


def stn3d_forward(sd, x, prefix="feat.stn.", training=False):
    """STN3d.forward -- model/pointnet.py:27-45."""
    B, _, N = x.shape
    h = F.relu(_bn(F.conv1d(x, sd[prefix + "conv1.weight"], sd[prefix + "conv1.bias"]), sd, prefix + "bn1", training))  # :29
    h = F.relu(_bn(F.conv1d(h, sd[prefix + "conv2.weight"], sd[prefix + "conv2.bias"]), sd, prefix + "bn2", training))  # :30
    h = F.relu(_bn(F.conv1d(h, sd[prefix + "conv3.weight"], sd[prefix + "conv3.bias"]), sd, prefix + "bn3", training))  # :31
    h = F.max_pool1d(h, N).view(-1, 1024)                                                                               # :32-33
    h = F.relu(_bn(F.linear(h, sd[prefix + "fc1.weight"], sd[prefix + "fc1.bias"]), sd, prefix + "bn4", training))      # :35
    h = F.relu(_bn(F.linear(h, sd[prefix + "fc2.weight"], sd[prefix + "fc2.bias"]), sd, prefix + "bn5", training))      # :36
    h = F.linear(h, sd[prefix + "fc3.weight"], sd[prefix + "fc3.bias"])                                                 # :37
    iden = torch.eye(3, dtype=h.dtype, device=h.device).reshape(1, 9)                                                   # :39-42
    return (h + iden).view(-1, 3, 3)                                                                                    # :43-44


def pointnetfeat_forward(sd, x, prefix="feat.", training=False):
    """PointNetfeat.forward (global_feat=True) -- model/pointnet.py:137-151."""
    B, _, N = x.shape
    trans = stn3d_forward(sd, x, prefix + "stn.", training)                                                             # :139
    h = torch.bmm(x.transpose(2, 1), trans).transpose(2, 1)                                                             # :140-143
    h = F.relu(_bn(F.conv1d(h, sd[prefix + "conv1.weight"], sd[prefix + "conv1.bias"]), sd, prefix + "bn1", training))  # :144
    h = F.relu(_bn(F.conv1d(h, sd[prefix + "conv2.weight"], sd[prefix + "conv2.bias"]), sd, prefix + "bn2", training))  # :146
    h = _bn(F.conv1d(h, sd[prefix + "conv3.weight"], sd[prefix + "conv3.bias"]), sd, prefix + "bn3", training)          # :147 (no ReLU)
    h = F.max_pool1d(h, N).view(-1, 1024)                                                                               # :148-149
    return h, trans


def pointnetcls_forward(sd, x, training=False):
    """PointNetCls.forward -- model/pointnet.py:189-194.  Returns (log_probs[B,k], trans[B,3,3])."""
    g, trans = pointnetfeat_forward(sd, x, "feat.", training)                                                           # :190
    h = F.relu(_bn(F.linear(g, sd["fc1.weight"], sd["fc1.bias"]), sd, "bn1", training))                                 # :191
    h = F.relu(_bn(F.linear(h, sd["fc2.weight"], sd["fc2.bias"]), sd, "bn2", training))                                 # :192
    h = F.linear(h, sd["fc3.weight"], sd["fc3.bias"])                                                                   # :193
    return F.log_softmax(h, dim=-1), trans                                                                              # :194


def to_torch_state(np_state, dtype=torch.float32, requires_grad=False):
    """numpy state dict (oracle.weights.make_state) -> torch tensors (buffers stay no-grad)."""
    sd = {}
    for k, v in np_state.items():
        if k.endswith("num_batches_tracked"):
            sd[k] = torch.tensor(int(v), dtype=torch.int64)
        else:
            t = torch.tensor(v, dtype=dtype)
            if requires_grad and not (k.endswith("running_mean") or k.endswith("running_var")):
                t.requires_grad_(True)
            sd[k] = t
    return sd


def train_step(sd, x, target, dlogp_extra=None):
    """One forward + nll_loss + backward, as main_1v.py:72-75.
    Returns (logp, trans, loss, grads dict)."""
    for k, v in sd.items():
        if v.requires_grad and v.grad is not None:
            v.grad = None
    logp, trans = pointnetcls_forward(sd, x, training=True)
    loss = F.nll_loss(logp, target)                     # main_1v.py:74
    loss.backward()                                     # main_1v.py:75
    grads = {k: v.grad.detach().clone() for k, v in sd.items() if v.requires_grad}
    return logp.detach(), trans.detach(), loss.detach(), grads
