"""CPU / eager-GPU port of the reference's dual-cloud network, functional form: SimpleSTN3d (PointNetGPD/model/pointnet.py:48-85),
DualPointNetfeat (:88-120) and DualPointNetCls (:157-174).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): the checker of pointnetgpd_b200's `pgpd_dual_*` path (csrc/dual.cuh).  Every
arithmetic step of the reference is a torch op; this restates the same sequence on a state dict (reference line numbers in the
comments).  Pinned to the imported, unmodified reference classes in tests/test_dual.py::test_dual_port_matches_reference (build
container only) and through the committed fixture tests/golden/dual_b6_n72_k2.npz (oracle/make_golden.py)."""
import torch
import torch.nn.functional as F

from .pointnet_torch_port import _bn


def simple_stn3d_forward(sd, x, prefix, training=False):
    """SimpleSTN3d.forward -- model/pointnet.py:68-85.  x: [B,3,N] -> [B,3,3]."""
    N = x.shape[2]
    h = F.relu(_bn(F.conv1d(x, sd[prefix + "conv1.weight"], sd[prefix + "conv1.bias"]), sd, prefix + "bn1", training))  # :70
    h = F.relu(_bn(F.conv1d(h, sd[prefix + "conv2.weight"], sd[prefix + "conv2.bias"]), sd, prefix + "bn2", training))  # :71
    h = F.relu(_bn(F.conv1d(h, sd[prefix + "conv3.weight"], sd[prefix + "conv3.bias"]), sd, prefix + "bn3", training))  # :72
    h = F.max_pool1d(h, N).view(-1, 256)                                                                                # :73-74
    h = F.relu(_bn(F.linear(h, sd[prefix + "fc1.weight"], sd[prefix + "fc1.bias"]), sd, prefix + "bn4", training))      # :76
    h = F.relu(_bn(F.linear(h, sd[prefix + "fc2.weight"], sd[prefix + "fc2.bias"]), sd, prefix + "bn5", training))      # :77
    h = F.linear(h, sd[prefix + "fc3.weight"], sd[prefix + "fc3.bias"])                                                 # :78
    iden = torch.eye(3, dtype=h.dtype, device=h.device).reshape(1, 9)                                                   # :80-83
    return (h + iden).view(-1, 3, 3)                                                                                    # :83-85


def dual_feat_forward(sd, x, prefix="feat.", training=False):
    """DualPointNetfeat.forward (global_feat=True) -- model/pointnet.py:103-117.  x: [B,6,N] -> ([B,1024], trans1 + trans2)."""
    N = x.shape[2]
    t1 = simple_stn3d_forward(sd, x[:, 0:3, :], prefix + "stn1.", training)                                             # :105
    t2 = simple_stn3d_forward(sd, x[:, 3:6, :], prefix + "stn2.", training)                                             # :106
    xt = x.transpose(2, 1)                                                                                              # :107
    h = torch.cat([torch.bmm(xt[..., 0:3], t1), torch.bmm(xt[..., 3:6], t2)], dim=-1).transpose(2, 1)                   # :108-109
    h = F.relu(_bn(F.conv1d(h, sd[prefix + "conv1.weight"], sd[prefix + "conv1.bias"]), sd, prefix + "bn1", training))  # :110
    h = F.relu(_bn(F.conv1d(h, sd[prefix + "conv2.weight"], sd[prefix + "conv2.bias"]), sd, prefix + "bn2", training))  # :112
    h = _bn(F.conv1d(h, sd[prefix + "conv3.weight"], sd[prefix + "conv3.bias"]), sd, prefix + "bn3", training)          # :113 (no ReLU)
    h = F.max_pool1d(h, N).view(-1, 1024)                                                                               # :114-115
    return h, t1 + t2                                                                                                   # :117


def dual_cls_forward(sd, x, training=False):
    """DualPointNetCls.forward -- model/pointnet.py:169-174.  Returns (log_probs [B,k], trans1 + trans2 [B,3,3])."""
    g, trans = dual_feat_forward(sd, x, "feat.", training)                                                              # :170
    h = F.relu(_bn(F.linear(g, sd["fc1.weight"], sd["fc1.bias"]), sd, "bn1", training))                                 # :171
    h = F.relu(_bn(F.linear(h, sd["fc2.weight"], sd["fc2.bias"]), sd, "bn2", training))                                 # :172
    h = F.linear(h, sd["fc3.weight"], sd["fc3.bias"])                                                                   # :173
    return F.log_softmax(h, dim=-1), trans                                                                              # :174
