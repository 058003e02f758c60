"""torch.autograd glue between the reference's nn.Module call surface and libpgpd's C ABI.

PyTorch is plumbing here: it owns device memory (caching allocator), the stream and the autograd
graph edge.  All arithmetic of the path happens inside libpgpd (hand-written sm_100a CUDA).
There is no PyTorch-op fallback: inputs the fused path does not cover raise.
"""
import ctypes as C

import torch

from . import _abi as A

_KEY_CACHE = {}

# Optional gradient-exchange hook of a data-parallel trainer (pointnetgpd_b200.ddp.FlatGradAllReduce.install):
#   hook(stage, flat, lo, hi) is called from inside the backward with stage 0 as soon as flat[lo:hi] (the gradients of the
#   classifier head + trunk tower) is final -- the T-Net half of the backward has not been launched yet, so an all-reduce issued
#   on a side stream overlaps it -- and with stage 1 once the rest (flat[lo:hi] of the second call) is final.
_GRAD_HOOK = None


def set_grad_hook(fn):
    """Install (or, with None, remove) the gradient-exchange hook; returns the previous one."""
    global _GRAD_HOOK
    old, _GRAD_HOOK = _GRAD_HOOK, fn
    return old


def _flat_layout(params, pkeys):
    """Offsets (in floats, 16-byte aligned) of every gradient inside ONE flat buffer, in ABI key order, and the boundary
    between the T-Net half (feat.stn.*: computed LAST by the backward) and the rest."""
    offs, off, split = [], 0, None
    for key, p in zip(pkeys, params):
        if split is None and not key.startswith("feat.stn."):
            split = off
        offs.append(off)
        off += (p.numel() + 3) // 4 * 4
    return offs, off, (off if split is None else split)


class _DeviceCtx:
    """current-device guard + raw stream handle for a CUDA device.  (For host tensors -- only ever
    reached by the unit tests that drive this glue through the SIMT emulator build of libpgpd --
    it is a no-op with the null stream.)"""

    def __init__(self, dev):
        self.dev = dev
        self.guard = torch.cuda.device(dev) if dev.type == "cuda" else None

    def __enter__(self):
        if self.guard is not None:
            self.guard.__enter__()
            return torch.cuda.current_stream(self.dev).cuda_stream
        return None

    def __exit__(self, *exc):
        if self.guard is not None:
            return self.guard.__exit__(*exc)
        return False


def _keys(what):
    if what not in _KEY_CACHE:
        _KEY_CACHE[what] = (tuple(A.param_keys(what)), tuple(A.buffer_keys(what)))
    return _KEY_CACHE[what]


# prefix of the ABI (PointNetCls-relative) key names to strip for stand-alone sub-modules
_STRIP = {A.PGPD_CLS: "", A.PGPD_FEAT: "feat.", A.PGPD_STN: "feat.stn."}


def _resolve(module, dotted):
    obj = module
    for part in dotted.split("."):
        obj = getattr(obj, part)
    return obj


def gather_tensors(module, what):
    """Parameters and BatchNorm buffers of `module` in ABI order."""
    pkeys, bkeys = _keys(what)
    strip = _STRIP[what]
    params = [_resolve(module, k[len(strip):]) for k in pkeys]
    bufs = [_resolve(module, k[len(strip):]) for k in bkeys]
    return params, bufs


def _check_tensor(t, name, device):
    if t.device != device:
        raise RuntimeError("pgpd: %s is on %s but the input is on %s" % (name, t.device, device))
    if t.dtype not in (torch.float32, torch.int64):
        raise TypeError("pgpd: %s must be float32 (got %s); the fused path computes in fp32" % (name, t.dtype))
    if not t.is_contiguous():
        t = t.contiguous()
    if t.data_ptr() % 16:
        # the ABI wants 16-byte aligned weights (vector loads).  nn.DataParallel replicas are views into a coalesced
        # broadcast buffer at arbitrary 4-byte offsets (torch.nn.parallel.replicate): take an aligned copy.
        t = t.clone(memory_format=torch.contiguous_format)
    return t


class _Fused(torch.autograd.Function):
    """forward(what, training, k, flags_extra, x, *params, *buffers) -> (out, trans)"""

    @staticmethod
    def forward(ctx, what, training, k, flags_extra, x, *tensors):
        lib = A.load()
        pkeys, bkeys = _keys(what)
        n_p = len(pkeys)
        dev = x.device
        params = [_check_tensor(t, pkeys[i], dev) for i, t in enumerate(tensors[:n_p])]
        bufs = list(tensors[n_p:])
        for i, b in enumerate(bufs):
            if b is None:
                raise RuntimeError("pgpd: BatchNorm without running statistics is not supported (%s)" % bkeys[i])
            if b.device != dev or not b.is_contiguous():
                raise RuntimeError("pgpd: buffer %s must be contiguous on %s" % (bkeys[i], dev))
        B, _, N = x.shape
        need_grad = any(ctx.needs_input_grad[5:5 + n_p])
        if ctx.needs_input_grad[4]:
            raise NotImplementedError("pgpd: gradient w.r.t. the input points is not provided "
                                      "(no reference script requests it: main_1v.py:69-75)")
        save = bool(need_grad and training)
        flags = (A.F_TRAIN if training else 0) | (A.F_SAVE if save else 0) | flags_extra
        table = dict(zip(pkeys, params))
        table.update(zip(bkeys, bufs))
        model = A.build_model(lambda key: table[key].data_ptr(), what)
        with _DeviceCtx(dev) as stream:
            nbytes = lib.pgpd_workspace_bytes(what, B, N, k, flags)
            ws = torch.empty(nbytes + 256, dtype=torch.uint8, device=dev)   # +256: the ABI wants 256-B alignment
            out = torch.empty((B, k if what == A.PGPD_CLS else 1024), dtype=torch.float32, device=dev)
            trans = torch.empty((B, 3, 3), dtype=torch.float32, device=dev)
            rc = lib.pgpd_forward(what, C.byref(model), x.data_ptr(), B, N, k, flags,
                                  out.data_ptr() if what != A.PGPD_STN else None, trans.data_ptr(),
                                  _aligned(ws), nbytes, stream)
        A.check(lib, rc)
        ctx.what, ctx.k, ctx.flags, ctx.training = what, k, flags, training
        ctx.saved = save
        if need_grad:
            ctx.save_for_backward(x, ws if save else None, *params, *bufs)
        if what == A.PGPD_STN:
            ctx.mark_non_differentiable(out)
        return out, trans

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dout, dtrans):
        if not ctx.saved:
            raise NotImplementedError("pgpd: backward through eval-mode BatchNorm is not implemented "
                                      "(call model.train() before a training step, as main_1v.py:62 does)")
        lib = A.load()
        what, k = ctx.what, ctx.k
        pkeys, bkeys = _keys(what)
        n_p = len(pkeys)
        x, ws = ctx.saved_tensors[0], ctx.saved_tensors[1]
        params = ctx.saved_tensors[2:2 + n_p]
        bufs = ctx.saved_tensors[2 + n_p:]
        dev = x.device
        B, _, N = x.shape
        table = dict(zip(pkeys, params))
        table.update(zip(bkeys, bufs))
        model = A.build_model(lambda key: table[key].data_ptr(), what)
        with _DeviceCtx(dev) as stream:
            # every gradient is a view into ONE flat buffer (ABI key order: T-Net tower, T-Net head | trunk, classifier head), so
            # that a data-parallel trainer can all-reduce it in place, in two buckets, without packing copies
            offs, total, split = _flat_layout(params, pkeys)
            flat = torch.zeros(total, dtype=torch.float32, device=dev)
            grads = [flat[o:o + p.numel()].view_as(p) for o, p in zip(offs, params)]
            gtable = dict(zip(pkeys, grads))
            g = A.build_grads(lambda key: gtable[key].data_ptr(), what)
            if what != A.PGPD_STN:
                dout = torch.zeros_like(ctx_out_like(B, k, what, dev)) if dout is None else dout.contiguous().float()
            if dtrans is not None:
                dtrans = dtrans.contiguous().float()
            elif what == A.PGPD_STN:
                dtrans = torch.zeros((B, 3, 3), dtype=torch.float32, device=dev)
            hook = _GRAD_HOOK if what != A.PGPD_STN else None

            def call(extra):
                return lib.pgpd_backward(what, C.byref(model), C.byref(g), x.data_ptr(), B, N, k, ctx.flags | extra,
                                         dout.data_ptr() if what != A.PGPD_STN else None,
                                         dtrans.data_ptr() if dtrans is not None else None,
                                         _aligned(ws), ws.numel() - 256, stream)
            if hook is None:
                rc = call(0)
            else:
                rc = call(A.F_BWD_HEAD)
                if rc == 0:
                    hook(0, flat, split, total)
                    rc = call(A.F_BWD_STN)
                    if rc == 0:
                        hook(1, flat, 0, split)
        A.check(lib, rc)
        return (None, None, None, None, None) + tuple(grads) + (None,) * len(bufs)


def _aligned(ws):
    p = ws.data_ptr()
    return p + ((-p) % 256)


def ctx_out_like(B, k, what, dev):
    return torch.empty((B, k if what == A.PGPD_CLS else 1024), dtype=torch.float32, device=dev)


def _prepare_input(x, num_points, input_chann):
    if not isinstance(x, torch.Tensor) or x.dim() != 3:
        raise ValueError("expected a [B, 3, N] tensor")
    if not x.is_cuda:
        raise RuntimeError("pointnetgpd_b200: the fused PointNet path is CUDA-only (input is on %s); "
                           "there is no CPU implementation in this package" % x.device)
    if input_chann != 3 or x.shape[1] != 3:
        raise ValueError("pointnetgpd_b200: only input_chann == 3 is supported (torch.bmm with the 3x3 "
                         "transform, pointnet.py:141, only works for 3 channels in the reference as well)")
    if x.shape[2] != num_points:
        raise ValueError("pointnetgpd_b200: got %d points per cloud but the model was built with num_points=%d "
                         "(MaxPool1d(num_points), pointnet.py:15,133, pools exactly num_points)" % (x.shape[2], num_points))
    if x.dtype != torch.float32:
        raise TypeError("pointnetgpd_b200: input must be float32 (main_1v.py:69 calls data.float())")
    return x if x.is_contiguous() else x.contiguous()


def _input_channels(module, what):
    """Input channels the module's first convolutions were built for (also correct for un-pickled reference checkpoints)."""
    convs = {A.PGPD_CLS: ("feat.conv1", "feat.stn.conv1"), A.PGPD_FEAT: ("conv1", "stn.conv1"), A.PGPD_STN: ("conv1",)}[what]
    chans = {int(_resolve(module, c).weight.shape[1]) for c in convs}
    return chans.pop() if len(chans) == 1 else -1


def run_module(module, what, x, k=1, flags_extra=0):
    """Evaluate STN3d / PointNetfeat / PointNetCls `module` on x through libpgpd."""
    num_points = module.num_points
    x = _prepare_input(x, num_points, _input_channels(module, what))
    params, bufs = gather_tensors(module, what)
    out, trans = _Fused.apply(what, bool(module.training), int(k), int(flags_extra), x, *params, *bufs)
    return out, trans


# ---- the dual-cloud network: SimpleSTN3d / DualPointNetfeat / DualPointNetCls (pointnet.py:48-120,157-174) -> pgpd_dual_* ----------
_DUAL_KEYS = {}


def _dual_keys(what):
    if what not in _DUAL_KEYS:
        _DUAL_KEYS[what] = (tuple(A.dual_param_keys(what)), tuple(A.dual_buffer_keys(what)))
    return _DUAL_KEYS[what]


class _DualFn(torch.autograd.Function):
    """forward(what, training, k, x, *params, *buffers) -> (out, trans)"""

    @staticmethod
    def forward(ctx, what, training, k, x, *tensors):
        lib = A.load()
        pkeys, bkeys = _dual_keys(what)
        n_p = len(pkeys)
        dev = x.device
        params = [_check_tensor(t, pkeys[i], dev) for i, t in enumerate(tensors[:n_p])]
        bufs = list(tensors[n_p:])
        for i, b in enumerate(bufs):
            if b is None or b.device != dev or not b.is_contiguous():
                raise RuntimeError("pgpd: buffer %s must be a contiguous tensor on %s" % (bkeys[i], dev))
        B, _, N = x.shape
        need_grad = any(ctx.needs_input_grad[4:4 + n_p])
        if ctx.needs_input_grad[3]:
            raise NotImplementedError("pgpd: gradient w.r.t. the input points is not provided")
        save = bool(need_grad and training)
        flags = (A.F_TRAIN if training else 0) | (A.F_SAVE if save else 0)
        table = dict(zip(pkeys, params))
        table.update(zip(bkeys, bufs))
        model = A.build_dual(lambda key: table[key].data_ptr(), what)
        width = {A.PGPD_DUAL_CLS: k, A.PGPD_DUAL_FEAT: 1024, A.PGPD_DUAL_STN: 1}[what]
        with _DeviceCtx(dev) as stream:
            nbytes = lib.pgpd_dual_workspace_bytes(what, B, N, k, flags)
            ws = torch.empty(nbytes + 256, dtype=torch.uint8, device=dev)
            out = torch.empty((B, width), dtype=torch.float32, device=dev)
            trans = torch.empty((B, 3, 3), dtype=torch.float32, device=dev)
            rc = lib.pgpd_dual_forward(what, C.byref(model), x.data_ptr(), B, N, k, flags, out.data_ptr(), trans.data_ptr(),
                                       _aligned(ws), nbytes, stream)
        A.check(lib, rc)
        ctx.what, ctx.k, ctx.flags, ctx.saved = what, k, flags, save
        if need_grad:
            ctx.save_for_backward(x, ws if save else None, *params, *bufs)
        if what == A.PGPD_DUAL_STN:
            ctx.mark_non_differentiable(out)
        return out, trans

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dout, dtrans):
        if not ctx.saved:
            raise NotImplementedError("pgpd: backward through eval-mode BatchNorm is not implemented (call model.train() first)")
        lib = A.load()
        what, k = ctx.what, ctx.k
        pkeys, bkeys = _dual_keys(what)
        n_p = len(pkeys)
        x, ws = ctx.saved_tensors[0], ctx.saved_tensors[1]
        params = ctx.saved_tensors[2:2 + n_p]
        bufs = ctx.saved_tensors[2 + n_p:]
        dev = x.device
        B, _, N = x.shape
        table = dict(zip(pkeys, params))
        table.update(zip(bkeys, bufs))
        model = A.build_dual(lambda key: table[key].data_ptr(), what)
        grads = [torch.empty_like(p) for p in params]
        gtable = dict(zip(pkeys, grads))
        g = A.build_dual(lambda key: gtable[key].data_ptr(), what, grad=True)
        with _DeviceCtx(dev) as stream:
            if what != A.PGPD_DUAL_STN:
                dout = torch.zeros((B, k if what == A.PGPD_DUAL_CLS else 1024), dtype=torch.float32, device=dev) if dout is None \
                    else dout.contiguous().float()
            if dtrans is not None:
                dtrans = dtrans.contiguous().float()
            elif what == A.PGPD_DUAL_STN:
                dtrans = torch.zeros((B, 3, 3), dtype=torch.float32, device=dev)
            rc = lib.pgpd_dual_backward(what, C.byref(model), C.byref(g), x.data_ptr(), B, N, k, ctx.flags,
                                        dout.data_ptr() if what != A.PGPD_DUAL_STN else None,
                                        dtrans.data_ptr() if dtrans is not None else None, _aligned(ws), ws.numel() - 256, stream)
        A.check(lib, rc)
        return (None, None, None, None) + tuple(grads) + (None,) * len(bufs)


def run_dual(module, what, x, k=1):
    """Evaluate SimpleSTN3d / DualPointNetfeat / DualPointNetCls `module` on x ([B,3,N] / [B,6,N]) through libpgpd."""
    chans = 3 if what == A.PGPD_DUAL_STN else 6
    if not isinstance(x, torch.Tensor) or x.dim() != 3 or x.shape[1] != chans:
        raise ValueError("expected a [B, %d, N] tensor" % chans)
    if not x.is_cuda:
        raise RuntimeError("pointnetgpd_b200: the dual-cloud path is CUDA-only (input is on %s); there is no CPU implementation in "
                           "this package" % x.device)
    if x.shape[2] != module.num_points:
        raise ValueError("pointnetgpd_b200: got %d points per cloud but the model was built with num_points=%d "
                         "(MaxPool1d(num_points), pointnet.py:54,99)" % (x.shape[2], module.num_points))
    if x.dtype != torch.float32:
        raise TypeError("pointnetgpd_b200: input must be float32")
    x = x if x.is_contiguous() else x.contiguous()
    pkeys, bkeys = _dual_keys(what)
    params = [_resolve(module, key) for key in pkeys]
    bufs = [_resolve(module, key) for key in bkeys]
    first = {A.PGPD_DUAL_STN: ("conv1", 3), A.PGPD_DUAL_FEAT: ("conv1", 6), A.PGPD_DUAL_CLS: ("feat.conv1", 6)}[what]
    if int(_resolve(module, first[0]).weight.shape[1]) != first[1]:
        raise ValueError("pointnetgpd_b200: the dual-cloud network needs input_chann=6 (3 per T-Net): torch.bmm with the 3x3 transforms "
                         "(pointnet.py:108) only works for two 3-channel halves in the reference as well")
    return _DualFn.apply(what, bool(module.training), int(k), x, *params, *bufs)
