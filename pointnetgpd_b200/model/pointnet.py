"""`model.pointnet` with the reference's call surface (PointNetGPD/model/pointnet.py).

STN3d (:8-45), PointNetfeat (:123-154) and PointNetCls (:177-194) keep the reference's
constructor signatures, attribute names and the 74 state_dict keys, so `torch.save(model)` pickles,
the shipped checkpoint and `nn.DataParallel` keep working -- but `forward` hands the whole
computation to libpgpd through `pointnetgpd_b200.functional` (one C call forward, one backward).

The sub-modules (`conv1`, `bn1`, `fc1`, `mp1`, `relu`, ...) exist only as parameter containers
with the reference's names; their own `forward` is never called.

SimpleSTN3d (:48-85), DualPointNetfeat (:88-120) and DualPointNetCls (:157-174) -- imported by main_1v.py:16 but
constructed by no reference script -- run through libpgpd's `pgpd_dual_*` entry points (fp32 CUDA-core kernels,
csrc/dual.cuh).  PointNetDenseCls (:197-221, per-point segmentation; no script, no caller) resolves as a name and
raises on construction.
"""
import torch.nn as nn

from .. import _abi as A
from ..functional import run_dual, run_module


def _shared_mlp(owner, input_chann, num_points):
    """conv1-3 / bn1-3 / mp1 with the reference's attribute names (pointnet.py:12-15,21-23)."""
    owner.conv1 = nn.Conv1d(input_chann, 64, 1)
    owner.conv2 = nn.Conv1d(64, 128, 1)
    owner.conv3 = nn.Conv1d(128, 1024, 1)
    owner.mp1 = nn.MaxPool1d(num_points)


class STN3d(nn.Module):
    """Input T-Net: [B,3,N] -> [B,3,3]  (pointnet.py:8-45)."""

    def __init__(self, num_points=2500, input_chann=3):
        super().__init__()
        self.num_points = num_points
        _shared_mlp(self, input_chann, num_points)
        self.fc1 = nn.Linear(1024, 512)
        self.fc2 = nn.Linear(512, 256)
        self.fc3 = nn.Linear(256, 9)
        self.relu = nn.ReLU()
        self.bn1 = nn.BatchNorm1d(64)
        self.bn2 = nn.BatchNorm1d(128)
        self.bn3 = nn.BatchNorm1d(1024)
        self.bn4 = nn.BatchNorm1d(512)
        self.bn5 = nn.BatchNorm1d(256)

    def forward(self, x):
        _, trans = run_module(self, A.PGPD_STN, x)
        return trans


class PointNetfeat(nn.Module):
    """T-Net + transform + trunk tower + max-pool: [B,3,N] -> ([B,1024], [B,3,3])  (pointnet.py:123-154)."""

    def __init__(self, num_points=2500, input_chann=3, global_feat=True):
        super().__init__()
        self.stn = STN3d(num_points=num_points, input_chann=input_chann)
        _shared_mlp(self, input_chann, num_points)
        # registration order of the reference: conv1-3, bn1-3, mp1
        mp1 = self._modules.pop("mp1")
        self.bn1 = nn.BatchNorm1d(64)
        self.bn2 = nn.BatchNorm1d(128)
        self.bn3 = nn.BatchNorm1d(1024)
        self.mp1 = mp1
        self.num_points = num_points
        self.global_feat = global_feat

    def forward(self, x):
        if not self.global_feat:
            raise NotImplementedError("PointNetfeat(global_feat=False) is used by no reference script "
                                      "(pointnet.py:152-154) and is outside this package's scope")
        return run_module(self, A.PGPD_FEAT, x)


class PointNetCls(nn.Module):
    """The grasp-quality classifier: [B,3,N] -> (log_probs [B,k], trans [B,3,3])  (pointnet.py:177-194)."""

    def __init__(self, num_points=2500, input_chann=3, k=2):
        super().__init__()
        self.num_points = num_points
        self.feat = PointNetfeat(num_points, input_chann=input_chann, global_feat=True)
        self.fc1 = nn.Linear(1024, 512)
        self.fc2 = nn.Linear(512, 256)
        self.fc3 = nn.Linear(256, k)
        self.bn1 = nn.BatchNorm1d(512)
        self.bn2 = nn.BatchNorm1d(256)
        self.relu = nn.ReLU()

    def forward(self, x):
        return run_module(self, A.PGPD_CLS, x, k=self.fc3.out_features)


class SimpleSTN3d(nn.Module):
    """The dual network's T-Net: [B,3,N] -> [B,3,3], tower 3->64->128->256, head 256->128->64->9  (pointnet.py:48-85)."""

    def __init__(self, num_points=2500, input_chann=3):
        super().__init__()
        self.num_points = num_points
        self.conv1 = nn.Conv1d(input_chann, 64, 1)
        self.conv2 = nn.Conv1d(64, 128, 1)
        self.conv3 = nn.Conv1d(128, 256, 1)
        self.mp1 = nn.MaxPool1d(num_points)
        self.fc1 = nn.Linear(256, 128)
        self.fc2 = nn.Linear(128, 64)
        self.fc3 = nn.Linear(64, 9)
        self.relu = nn.ReLU()
        self.bn1 = nn.BatchNorm1d(64)
        self.bn2 = nn.BatchNorm1d(128)
        self.bn3 = nn.BatchNorm1d(256)
        self.bn4 = nn.BatchNorm1d(128)
        self.bn5 = nn.BatchNorm1d(64)

    def forward(self, x):
        _, trans = run_dual(self, A.PGPD_DUAL_STN, x)
        return trans


class DualPointNetfeat(nn.Module):
    """Two clouds per grasp: [B,6,N] -> ([B,1024], trans1 + trans2); one SimpleSTN3d per 3-channel half, 6-channel trunk
    (pointnet.py:88-120)."""

    def __init__(self, num_points=2500, input_chann=6, global_feat=True):
        super().__init__()
        self.stn1 = SimpleSTN3d(num_points=num_points, input_chann=input_chann // 2)
        self.stn2 = SimpleSTN3d(num_points=num_points, input_chann=input_chann // 2)
        self.conv1 = nn.Conv1d(input_chann, 64, 1)
        self.conv2 = nn.Conv1d(64, 128, 1)
        self.conv3 = nn.Conv1d(128, 1024, 1)
        self.bn1 = nn.BatchNorm1d(64)
        self.bn2 = nn.BatchNorm1d(128)
        self.bn3 = nn.BatchNorm1d(1024)
        self.mp1 = nn.MaxPool1d(num_points)
        self.num_points = num_points
        self.global_feat = global_feat

    def forward(self, x):
        if not self.global_feat:
            raise NotImplementedError("DualPointNetfeat(global_feat=False) is used by no reference script "
                                      "(pointnet.py:118-120) and is outside this package's scope")
        return run_dual(self, A.PGPD_DUAL_FEAT, x)


class DualPointNetCls(nn.Module):
    """[B,6,N] -> (log_probs [B,k], trans1 + trans2 [B,3,3])  (pointnet.py:157-174).  Like the reference's, the default
    input_chann=3 constructs a model whose forward cannot run (1-channel T-Nets against 3-channel slices, pointnet.py:105): pass
    input_chann=6."""

    def __init__(self, num_points=2500, input_chann=3, k=2):
        super().__init__()
        self.num_points = num_points
        self.feat = DualPointNetfeat(num_points, input_chann=input_chann, global_feat=True)
        self.fc1 = nn.Linear(1024, 512)
        self.fc2 = nn.Linear(512, 256)
        self.fc3 = nn.Linear(256, k)
        self.bn1 = nn.BatchNorm1d(512)
        self.bn2 = nn.BatchNorm1d(256)
        self.relu = nn.ReLU()

    def forward(self, x):
        return run_dual(self, A.PGPD_DUAL_CLS, x, k=self.fc3.out_features)


def _out_of_scope(name, where):
    def __init__(self, *args, **kwargs):
        raise NotImplementedError(
            "%s (%s) is constructed by no reference script and is outside the scope of pointnetgpd_b200 "
            "(SURVEY.md section 8f); the name exists only so that reference imports resolve." % (name, where))
    return type(name, (nn.Module,), {"__init__": __init__, "__module__": __name__})


PointNetDenseCls = _out_of_scope("PointNetDenseCls", "pointnet.py:197-221")

# BASELINE.json names a 3-class variant "PointNetClsMC"; in the reference it is PointNetCls(k=3)
# (main_1v_mc.py:103).  Provided as a convenience alias.
def PointNetClsMC(num_points=2500, input_chann=3):
    return PointNetCls(num_points=num_points, input_chann=input_chann, k=3)
