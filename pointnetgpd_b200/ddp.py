"""Data-parallel gradient exchange for the PointNet training step.

Grasp batches are independent, so training is plain data parallel: every rank holds a full model
replica and its own clouds; BatchNorm statistics stay per-rank (exactly what the reference's
nn.DataParallel does, main_1v.py:163-165); the only exchange per step is an all-reduce (average) of
the 1.6 M fp32 gradients (6.4 MB), over NCCL / NVLink on the GPU box (gloo in the CPU tests).

The fused backward writes every gradient as a view into ONE flat buffer and runs in two halves
(pgpd.h: PGPD_F_BWD_HEAD / PGPD_F_BWD_STN).  `FlatGradAllReduce.install()` hooks in between: the bucket
of the first half (classifier head + trunk tower, 3.2 MB) is all-reduced IN PLACE on a side stream
while the T-Net half of the backward is still being computed; the second bucket follows it.  No packing
copies, no separate scaling pass (ReduceOp.AVG), and -- all launches being asynchronous -- the whole
step including the collectives can be captured into one CUDA graph.
"""
import torch
import torch.distributed as dist


class FlatGradAllReduce:
    """all_reduce(): the exchange for gradients that were NOT produced through the hook (e.g. the CPU / gloo tests, or
    a model wrapped differently): packs into a persistent flat buffer, one all-reduce, writes the averages back in
    place.  install(): the overlapped, copy-free path described in the module docstring."""

    def __init__(self, params, world_size=None, group=None, overlap=True):
        self.overlap = overlap
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        if world_size is None:
            world_size = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.world = int(world_size)
        self.numel = sum(p.numel() for p in self.params)
        self.flat = None
        self.views = None
        self.side = None
        self.hooked_steps = 0
        self._in_hook_step = False

    # ------------------------------------------------------------------ overlapped path
    def install(self):
        """Route the gradient exchange through the fused backward's hook (pointnetgpd_b200.functional.set_grad_hook)."""
        from . import functional
        if self.world > 1:
            functional.set_grad_hook(self._hook)
        return self

    def uninstall(self):
        from . import functional
        functional.set_grad_hook(None)

    def _avg(self, t):
        if dist.get_backend(self.group) == "nccl":
            dist.all_reduce(t, op=dist.ReduceOp.AVG, group=self.group)
        else:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
            t.mul_(1.0 / self.world)

    def _hook(self, stage, flat, lo, hi):
        cur = torch.cuda.current_stream(flat.device)
        if self.side is None:
            self.side = torch.cuda.Stream(device=flat.device)
        if not self.overlap:
            # one all-reduce of the whole flat buffer on the compute stream once the backward is complete
            if stage == 1:
                self._avg(flat)
                self._in_hook_step = True
                self.hooked_steps += 1
            return
        if stage == 0:
            self.side.wait_stream(cur)                 # bucket 0 is final on the compute stream
            with torch.cuda.stream(self.side):
                self._avg(flat[lo:hi])
        else:
            self.side.wait_stream(cur)
            with torch.cuda.stream(self.side):
                self._avg(flat[lo:hi])
            cur.wait_stream(self.side)                 # join: the optimizer sees both buckets averaged
            self._in_hook_step = True
            self.hooked_steps += 1

    # ------------------------------------------------------------------ plain path
    def _ensure(self):
        if self.flat is None:
            p0 = self.params[0]
            self.flat = torch.zeros(self.numel, dtype=p0.dtype, device=p0.device)
            self.views, off = [], 0
            for p in self.params:
                self.views.append(self.flat[off:off + p.numel()].view_as(p))
                off += p.numel()

    def all_reduce(self):
        if self.world <= 1:
            return
        if self._in_hook_step:                         # already exchanged inside the backward
            self._in_hook_step = False
            return
        self._ensure()
        grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in self.params]
        torch._foreach_copy_(self.views, grads)
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
        self.flat.mul_(1.0 / self.world)
        for p, g in zip(self.params, grads):
            if p.grad is None:
                p.grad = g
        torch._foreach_copy_(grads, self.views)

    @property
    def bytes_per_step(self):
        return self.numel * 4


def broadcast_module_state(module, src=0, group=None):
    """Make every rank start from rank `src`'s parameters and buffers (DDP's initial sync)."""
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src=src, group=group)
