"""CUDA-graph capture of a whole training step of ``mae_b200.modules.MAE``.

The reference's training loop (experiments/color_image.py:153-165, experiments/binary_image.py:150-160) is

    optimizer.zero_grad(); loss, ... = mae.loss(data, nsamples, eta, gamma, free_bits); loss.backward()
    clip_grad_norm_(...); optimizer.step()

Eagerly, the step is host-bound on this path: every node of ours is one ctypes call plus an autograd node (~50 us of
Python per launch against kernels of a few microseconds), and the PyTorch conv stacks add their own launches.  With static
shapes the whole forward + backward (optionally clipping and the optimizer) replays as ONE graph launch:

    step = capture_step(model, x_example, nsamples, eta, gamma, free_bits, optimizer=opt, max_norm=5.0)
    for data in batches:
        loss, recon, kl, pkl_m, pkl_s, l_mean, l_std = step(data)      # device tensors, valid until the next call

Noise (posterior samples, dropout) comes from torch's graph-safe Philox generator, so every replay draws fresh samples.
The kernels of ours record their side-stream fork / join inside the capture (ops.py), so the regulariser still overlaps
with the decoder in the replayed graph.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import ops


class CapturedStep:
    """One captured ``loss -> backward [-> clip -> optimizer.step]`` of a model; call it with a batch of the captured shape."""

    def __init__(self, model, x_example: torch.Tensor, nsamples: int, eta: float = 0.0, gamma: float = 0.0,
                 free_bits: float = 0.0, optimizer: Optional[torch.optim.Optimizer] = None, max_norm: Optional[float] = None,
                 warmup: int = 3):
        if not x_example.is_cuda:
            raise RuntimeError("capture_step needs CUDA tensors (there is no CPU path)")
        self.model, self.optimizer, self.max_norm = model, optimizer, max_norm
        self.args = (int(nsamples), float(eta), float(gamma), float(free_bits))
        self.x = x_example.detach().clone()
        dev = self.x.device
        # the gradient handed to the loss is created by the step itself and is exactly one: the fused likelihood node then
        # needs no rescale launch in its backward (ops.set_unit_loss_grad)
        self._one = torch.ones((), device=dev)
        self.params = [p for p in model.parameters() if p.requires_grad]
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        prev_unit = ops.set_unit_loss_grad(True)
        try:
            with torch.cuda.stream(side):
                for _ in range(max(1, warmup)):          # cuDNN autotuning, workspace pools, lazy initialisation
                    self._zero_grads()
                    self._step_body(optimizer_step=False)
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize(dev)
            self._zero_grads()
            self.graph = torch.cuda.CUDAGraph()
            n0 = ops.launch_count()
            with torch.cuda.graph(self.graph):
                self.outputs = self._step_body(optimizer_step=True)
            self.launches_of_ours = ops.launch_count() - n0
            # the gradient tensors the captured backward writes into: a replay refreshes them in place; they are re-attached
            # after every call because ``optimizer.zero_grad()`` / ``model.zero_grad()`` set ``p.grad`` to None by default
            self.static_grads = [p.grad for p in self.params]
        finally:
            ops.set_unit_loss_grad(prev_unit)

    def _zero_grads(self):
        for p in self.params:
            p.grad = None

    def _step_body(self, optimizer_step: bool):
        ns, eta, gamma, free_bits = self.args
        for p in self.params:            # inside the capture: gradients are (re)created by the graph's own allocations
            p.grad = None
        out = self.model.loss(self.x, ns, eta=eta, gamma=gamma, free_bits=free_bits)
        out[0].backward(gradient=self._one)
        if self.max_norm is not None:
            torch.nn.utils.clip_grad_norm_(self.params, self.max_norm, foreach=True)
        if optimizer_step and self.optimizer is not None:
            self.optimizer.step()
        return tuple(o.detach() if isinstance(o, torch.Tensor) else o for o in out)

    def __call__(self, x: torch.Tensor):
        if x.shape != self.x.shape:
            raise RuntimeError("captured step expects a batch of shape %s, got %s" % (tuple(self.x.shape), tuple(x.shape)))
        self.x.copy_(x, non_blocking=True)
        self.graph.replay()
        for p, g in zip(self.params, self.static_grads):
            p.grad = g
        return self.outputs


def capture_step(model, x_example: torch.Tensor, nsamples: int, eta: float = 0.0, gamma: float = 0.0, free_bits: float = 0.0,
                 optimizer: Optional[torch.optim.Optimizer] = None, max_norm: Optional[float] = None, warmup: int = 3
                 ) -> CapturedStep:
    """Capture ``model.loss(x, nsamples, eta, gamma, free_bits)[0].backward()`` (and, when given, gradient clipping to
    ``max_norm`` and ``optimizer.step()``) for batches shaped like ``x_example``.  The optimizer should be constructed with
    ``capturable=True`` where PyTorch requires it (Adam / Adamax families).  Parameter gradients live in ``p.grad`` after
    each call, as in the eager loop."""
    return CapturedStep(model, x_example, nsamples, eta, gamma, free_bits, optimizer, max_norm, warmup)
