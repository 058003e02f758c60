"""`model.gpd` with the reference's call surface (PointNetGPD/model/gpd.py:5-31): `GPDClassifier(input_chann, dropout=False)`,
the paper's baseline CNN on 60 x 60 projections (constructed by main_1v_gpd.py:105 / main_fullv_gpd.py).  Same sub-module
names (`conv1`, `pool1`, `conv2`, `pool2`, `fc1`, `dp`, `relu`, `fc2`), so state_dicts and whole-module pickles round-trip;
`forward` hands the computation to libpgpd (`pgpd_gpd_forward` / `pgpd_gpd_backward`, csrc/gpd.cuh) -- there is no
PyTorch-op fallback.  Not provided: `dropout=True` in training mode (nn.Dropout2d draws from torch's RNG stream; raises),
gradients w.r.t. the input images, CPU tensors."""
import ctypes as C

import torch
import torch.nn as nn

from .. import _abi as A
from ..functional import _DeviceCtx, _aligned

__all__ = ["GPDClassifier"]


class _GpdFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, *params):
        lib = A.load()
        dev = x.device
        ps = []
        for p in params:
            if p.device != dev or p.dtype != torch.float32:
                raise RuntimeError("pgpd: GPDClassifier parameters must be float32 on %s" % dev)
            p = p.contiguous()
            if p.data_ptr() % 16:
                p = p.clone(memory_format=torch.contiguous_format)
            ps.append(p)
        B, Cc = int(x.shape[0]), int(x.shape[1])
        need_grad = any(ctx.needs_input_grad[1:])
        if ctx.needs_input_grad[0]:
            raise NotImplementedError("pgpd: gradient w.r.t. the input images is not provided")
        flags = A.F_SAVE if need_grad else 0
        m = A.Gpd()
        for i, name in enumerate(A.GPD_LAYERS):
            lin = getattr(m, name)
            lin.w, lin.b = ps[2 * i].data_ptr(), ps[2 * i + 1].data_ptr()
        with _DeviceCtx(dev) as stream:
            nbytes = lib.pgpd_gpd_workspace_bytes(B, Cc, flags)
            ws = torch.empty(nbytes + 256, dtype=torch.uint8, device=dev)
            logp = torch.empty((B, 2), dtype=torch.float32, device=dev)
            rc = lib.pgpd_gpd_forward(C.byref(m), x.data_ptr(), B, Cc, flags, logp.data_ptr(), _aligned(ws), nbytes, stream)
        A.check(lib, rc)
        ctx.flags = flags
        if need_grad:
            ctx.save_for_backward(x, ws, *ps)
        return logp

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dlogp):
        lib = A.load()
        x, ws = ctx.saved_tensors[0], ctx.saved_tensors[1]
        ps = ctx.saved_tensors[2:]
        dev = x.device
        B, Cc = int(x.shape[0]), int(x.shape[1])
        m, g = A.Gpd(), A.GpdGrad()
        grads = [torch.empty_like(p) for p in ps]
        for i, name in enumerate(A.GPD_LAYERS):
            lin, lg = getattr(m, name), getattr(g, name)
            lin.w, lin.b = ps[2 * i].data_ptr(), ps[2 * i + 1].data_ptr()
            lg.dw, lg.db = grads[2 * i].data_ptr(), grads[2 * i + 1].data_ptr()
        dlogp = dlogp.contiguous().float()
        with _DeviceCtx(dev) as stream:
            rc = lib.pgpd_gpd_backward(C.byref(m), C.byref(g), x.data_ptr(), B, Cc, ctx.flags, dlogp.data_ptr(), _aligned(ws),
                                       ws.numel() - 256, stream)
        A.check(lib, rc)
        return (None,) + tuple(grads)


class GPDClassifier(nn.Module):
    """Input: (batch_size, input_chann, 60, 60) -> log-probabilities (batch_size, 2)   (gpd.py:5-31)."""

    def __init__(self, input_chann, dropout=False):
        super().__init__()
        self.conv1 = nn.Conv2d(input_chann, 20, 5)
        self.pool1 = nn.MaxPool2d(2, stride=2)
        self.conv2 = nn.Conv2d(20, 50, 5)
        self.pool2 = nn.MaxPool2d(2, stride=2)
        self.fc1 = nn.Linear(12 * 12 * 50, 500)
        self.dp = nn.Dropout2d(p=0.5, inplace=False)
        self.relu = nn.ReLU()
        self.fc2 = nn.Linear(500, 2)
        self.if_dropout = dropout

    def forward(self, x):
        if not isinstance(x, torch.Tensor) or x.dim() != 4 or x.shape[2] != 60 or x.shape[3] != 60:
            raise ValueError("GPDClassifier expects a [B, input_chann, 60, 60] tensor (gpd.py:7)")
        if x.shape[1] != self.conv1.weight.shape[1]:
            raise ValueError("got %d input channels, the model was built for %d" % (x.shape[1], self.conv1.weight.shape[1]))
        if not x.is_cuda:
            raise RuntimeError("pointnetgpd_b200: GPDClassifier is CUDA-only (input is on %s)" % x.device)
        if x.dtype != torch.float32:
            raise TypeError("input must be float32 (main_1v_gpd.py calls data.float())")
        if self.if_dropout and self.training:
            raise NotImplementedError("GPDClassifier(dropout=True) in training mode (nn.Dropout2d, gpd.py:27-28) is not provided: the "
                                      "mask comes from torch's RNG stream; no reference script passes dropout=True")
        x = x if x.is_contiguous() else x.contiguous()
        return _GpdFn.apply(x, self.conv1.weight, self.conv1.bias, self.conv2.weight, self.conv2.bias,
                            self.fc1.weight, self.fc1.bias, self.fc2.weight, self.fc2.bias)
