"""Whole-step CUDA-graph replay for training: forward + loss + backward + optimizer step captured once, replayed per iteration.

A BEM training step (basicsr/models/image_restoration_model.py optimize_parameters: net_g(lq) -> losses -> backward -> clip ->
optimizer.step) is ~2000 kernels of 5-50 us at 8 x 128 x 128: launched eagerly it is bound by the host (44.7 ms per step on a B200
with either scan backend, the GPU idle most of the time). The kernels of this package never synchronise with the host and take
their stream from torch, so the whole step can be captured; the replay runs at the device's pace.

    step = GraphedTrainStep(model, loss_fn, optimizer, example_inputs)     # warms up, captures
    loss = step(*inputs)                                                    # copies the inputs into the static buffers, replays

Constraints (torch's for any whole-network capture): static shapes; an optimizer constructed with `capturable=True`; no host
synchronisation inside the step (`.item()`, data-dependent control flow). The eager step stays available as `step.eager(*inputs)`.
Under DistributedDataParallel (torch's rules for capturing the NCCL all-reduce with the step): set
TORCH_NCCL_ASYNC_ERROR_HANDLING=0 before init_process_group, construct the DDP wrapper inside `with torch.cuda.stream(side)` and
pass the same `stream=side` here (the reducer records the stream it was built on; the legacy default stream cannot take part in
a capture), and give it `warmup >= 11` iterations.
"""
from __future__ import annotations

import torch
from torch.overrides import TorchFunctionMode


class _DeviceIndexMode(TorchFunctionMode):
    """`t[:, [0, 2, 4]]` builds its index tensor on the host and copies it to the device on every call, which a stream capture
    refuses (the reference's decomposition front end does this: DecompDualBranchDDWavelet_arch.py:129-130). Under this mode a
    Python-list index is replaced by a device tensor made once, during the eager warm-up, and reused by the capture."""

    def __init__(self):
        super().__init__()
        self.cache = {}

    def _dev(self, idx, device):
        key = (tuple(idx), device)
        if key not in self.cache:
            self.cache[key] = torch.tensor(idx, dtype=torch.long, device=device)
        return self.cache[key]

    def __torch_function__(self, func, types, args=(), kwargs=None):
        if func is torch.Tensor.__getitem__ and len(args) == 2 and isinstance(args[0], torch.Tensor) and args[0].is_cuda:
            t, idx = args
            flat = lambda i: isinstance(i, list) and len(i) > 0 and all(isinstance(v, int) for v in i)
            if flat(idx):
                args = (t, self._dev(idx, t.device))
            elif isinstance(idx, tuple) and any(flat(i) for i in idx):
                args = (t, tuple(self._dev(i, t.device) if flat(i) else i for i in idx))
        return func(*args, **(kwargs or {}))


class GraphedTrainStep:
    def __init__(self, model, loss_fn, optimizer, example_inputs, warmup: int = 3, max_grad_norm=None, stream=None):
        self.model, self.loss_fn, self.opt = model, loss_fn, optimizer
        self.max_grad_norm = max_grad_norm
        self.static_in = [t.clone() for t in example_inputs]
        dev = self.static_in[0].device
        self.stream = stream if stream is not None else torch.cuda.Stream(device=dev)
        self.stream.wait_stream(torch.cuda.current_stream(dev))
        self._mode = _DeviceIndexMode()
        with torch.cuda.stream(self.stream), self._mode:          # warm-up on the capturing stream: allocator, workspaces, optimizer state
            for _ in range(max(1, warmup)):
                self._body()
        torch.cuda.current_stream(dev).wait_stream(self.stream)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, stream=self.stream), self._mode:
            self.static_loss = self._body()

    def _body(self):
        self.opt.zero_grad(set_to_none=False)          # gradients accumulate into the same tensors on every replay
        loss = self.loss_fn(self.model, *self.static_in)
        loss.backward()
        if self.max_grad_norm is not None:
            torch.nn.utils.clip_grad_norm_(self.model.parameters(), self.max_grad_norm, foreach=True)
        self.opt.step()
        return loss

    def eager(self, *inputs):
        for s, t in zip(self.static_in, inputs):
            s.copy_(t, non_blocking=True)
        return self._body()

    def __call__(self, *inputs):
        for s, t in zip(self.static_in, inputs):
            s.copy_(t, non_blocking=True)
        self.graph.replay()
        return self.static_loss
