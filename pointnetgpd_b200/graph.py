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
    """One graph: zero_grad, forward, loss, backward (with the gradient all-reduce issued from inside the backward on a side
    stream when `grad_sync` is hooked in, ddp.FlatGradAllReduce.install), optimizer.  `capture_sync=False` keeps collectives
    out of the graph: two graphs -- (forward, loss, backward) and (optimizer) -- with an eager all-reduce in between.

    NOTE: construction runs `warmup` REAL training steps on the example batch before capturing (CUDA-graph capture needs warmed-up
    allocator pools and library state).  To keep the caller's training trajectory untouched, parameters, BatchNorm buffers and the
    optimizer state are snapshotted before and restored after the warm-up."""

    def __init__(self, model, optimizer, x_example, y_example, grad_sync=None, warmup=3, loss_fn=F.nll_loss,
                 before_capture=None, capture_sync=True):
        self.model, self.opt, self.sync, self.loss_fn = model, optimizer, grad_sync, loss_fn
        dev = x_example.device if x_example.is_cuda else next(model.parameters()).device
        self.sx = torch.empty_like(x_example, device=dev)
        self.sy = torch.empty_like(y_example, device=dev)
        self.sx.copy_(x_example)
        self.sy.copy_(y_example)
        import copy
        snap_model = copy.deepcopy(model.state_dict())
        snap_opt = copy.deepcopy(optimizer.state_dict())
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._fwd_bwd()
                if self.sync is not None:
                    self.sync.all_reduce()
                self.opt.step()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        with torch.no_grad():                       # restore IN PLACE: the captured graph keeps the tensors' addresses
            for k, v in model.state_dict().items():
                v.copy_(snap_model[k])
        cur_opt = optimizer.state_dict()
        for idx, st in cur_opt["state"].items():
            for k, v in st.items():
                if torch.is_tensor(v):
                    if idx in snap_opt["state"] and k in snap_opt["state"][idx]:
                        v.copy_(snap_opt["state"][idx][k])
                    else:
                        v.zero_()                   # state created by the warm-up (Adam moments, step counter): back to its initial value
        torch.cuda.synchronize(dev)
        if before_capture is not None:
            before_capture()
        self.g_main = torch.cuda.CUDAGraph()
        self.g_opt = None
        # gradients keep the tensors of the last warm-up step: capture re-creates them inside the graph's pool
        self.opt.zero_grad(set_to_none=True)
        if self.sync is None or capture_sync:
            with torch.cuda.graph(self.g_main):
                self.loss = self._fwd_bwd()
                if self.sync is not None:
                    self.sync.all_reduce()          # no-op when the exchange already happened inside the backward
                self.opt.step()
        else:
            with torch.cuda.graph(self.g_main):
                self.loss = self._fwd_bwd()
            self.g_opt = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.g_opt, pool=self.g_main.pool()):
                self.opt.step()

    def _fwd_bwd(self):
        self.opt.zero_grad(set_to_none=True)
        logp, _ = self.model(self.sx)
        loss = self.loss_fn(logp, self.sy)
        loss.backward()
        return loss.detach()

    def staged_input(self):
        """A `staging.StagedInput` over this step's static inputs: `prefetch((x_host, y_host))` the next batch while a step runs,
        then `step_staged()`."""
        from .staging import StagedInput
        if getattr(self, "_staged", None) is None:
            self._staged = StagedInput([self.sx, self.sy])
        return self._staged

    def step_staged(self):
        """Replay on the batch last handed to `staged_input().prefetch`.  Returns the (static) loss tensor."""
        self._staged.commit()
        return self._replay()

    def step(self, x, y):
        """x, y: device tensors or pinned host tensors of the captured shapes.  Returns the (static) loss tensor."""
        self.sx.copy_(x, non_blocking=True)
        self.sy.copy_(y, non_blocking=True)
        return self._replay()

    def _replay(self):
        self.g_main.replay()
        if self.g_opt is not None:
            self.sync.all_reduce()          # eager NCCL all-reduce, in place on the graph's static gradient tensors
            self.g_opt.replay()
        return self.loss
