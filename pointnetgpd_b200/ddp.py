"""Data-parallel gradient exchange for the PointNet training step.

Grasp batches are independent, so training is plain data parallel: every rank holds a full model
replica and its own clouds; BatchNorm statistics stay per-rank (exactly what the reference's
nn.DataParallel does, main_1v.py:163-165); the only exchange per step is ONE all-reduce(sum)/world
of the 1.6 M fp32 gradients (6.4 MB), over NCCL / NVLink on the GPU box (gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


class FlatGradAllReduce:
    """Keeps one persistent flat fp32 buffer; after backward, packs all gradients into it, all-reduces it once and
    writes the averages back IN PLACE into the existing `p.grad` tensors (their addresses stay fixed, which is what a
    CUDA-graph captured optimizer step needs)."""

    def __init__(self, params, world_size=None, group=None):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        if world_size is None:
            world_size = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.world = int(world_size)
        self.numel = sum(p.numel() for p in self.params)
        self.flat = None
        self.views = None

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
