"""CUDA-graph capture of the training step.

The step launches ~500 small-to-medium kernels (libpgpd) plus the optimizer; at ~5 ms per step the host-side launch
path is a measurable fraction.  `GraphedTrainStep` captures forward + loss + backward (+ gradient all-reduce) +
optimizer step once and replays it; inputs are copied into static device buffers first (pinned host tensors copy
asynchronously on the same stream).  libpgpd never allocates or synchronises, so its launches are capturable as is;
the workspace / gradient tensors allocated by the autograd glue come from the graph's private memory pool.
"""
import torch
import torch.nn.functional as F


class GraphedTrainStep:
    def __init__(self, model, optimizer, x_example, y_example, grad_sync=None, warmup=3, loss_fn=F.nll_loss,
                 before_capture=None):
        self.model, self.opt, self.sync, self.loss_fn = model, optimizer, grad_sync, loss_fn
        self.sx = torch.empty_like(x_example, device=x_example.device if x_example.is_cuda else next(model.parameters()).device)
        self.sy = torch.empty_like(y_example, device=self.sx.device)
        self.sx.copy_(x_example)
        self.sy.copy_(y_example)
        side = torch.cuda.Stream(device=self.sx.device)
        side.wait_stream(torch.cuda.current_stream(self.sx.device))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._eager()
        torch.cuda.current_stream(self.sx.device).wait_stream(side)
        torch.cuda.synchronize(self.sx.device)
        if before_capture is not None:
            before_capture()
        self.graph = torch.cuda.CUDAGraph()
        self.opt.zero_grad(set_to_none=True)
        with torch.cuda.graph(self.graph):
            self.loss = self._eager()

    def _eager(self):
        self.opt.zero_grad(set_to_none=True)
        logp, _ = self.model(self.sx)
        loss = self.loss_fn(logp, self.sy)
        loss.backward()
        if self.sync is not None:
            self.sync.all_reduce()
        self.opt.step()
        return loss.detach()

    def step(self, x, y):
        """x, y: device tensors or pinned host tensors of the captured shapes.  Returns the (static) loss tensor."""
        self.sx.copy_(x, non_blocking=True)
        self.sy.copy_(y, non_blocking=True)
        self.graph.replay()
        return self.loss
