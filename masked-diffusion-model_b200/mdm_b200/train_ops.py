"""Fused training-step pieces used by the trainer drop-ins.

* `residual_mse`: `recon = (x_in + net) - shift; loss = mean(w_b (recon - x0)^2)` and its gradient
  w.r.t. `net` in ONE kernel (csrc/nn_kernels.cu: mse_residual_kernel), exposed as an autograd
  function so `accelerator.backward(loss)` -- what the reference calls (trainer_masked.py:126-142,
  trainer_masked_mean_shift.py:142-168) -- reaches the hand-written denoiser backward.
* `GraphedCallable`: warm-up + CUDA-graph capture + replay of a stream-ordered step.
"""
from __future__ import annotations

import torch

from . import denoiser_ops as ops


class _ResidualMSE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, net, x_in, shift, x0, weight):
        net = net.float().contiguous()
        x_in = x_in.float().contiguous()
        x0 = x0.float().contiguous()
        if shift is not None:
            shift = shift.float().expand_as(x_in).contiguous()
        if weight is not None:
            weight = weight.float().contiguous()
        dev = net.device
        dnet = torch.empty_like(net)
        recon = torch.empty_like(net)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        ws = torch.empty(1024, dtype=torch.float32, device=dev)
        per_sample = net.numel() // net.shape[0]
        ops.mse_residual(x_in, net, shift, x0, weight, dnet, recon, loss, ws, per_sample)
        ctx.save_for_backward(dnet)
        ctx.mark_non_differentiable(recon)
        return loss, recon

    @staticmethod
    def backward(ctx, g_loss, _g_recon):
        (dnet,) = ctx.saved_tensors
        return dnet * g_loss, None, None, None, None


def residual_mse(net, x_in, x0, shift=None, weight=None):
    """-> (loss, recon): loss = mean(weight[b] * ((x_in + net) - shift - x0)^2), recon = (x_in + net) - shift"""
    if not net.is_cuda:
        raise RuntimeError("residual_mse: CUDA tensors only (no CPU fallback)")
    return _ResidualMSE.apply(net, x_in, shift, x0, weight)


class GraphedCallable:
    """Runs `fn()` eagerly `warmup` times, then captures it into a CUDA graph and replays it.

    `fn` must be stream-ordered with no host synchronisation, read its inputs from fixed device
    buffers and return a (nest of) tensor(s) that stay valid between replays."""

    def __init__(self, fn, warmup=2, enabled=True):
        self.fn, self.warmup, self.enabled = fn, warmup, enabled
        self.calls = 0
        self.graph = None
        self.result = None

    def __call__(self):
        if not self.enabled:
            return self.fn()
        if self.graph is not None:
            self.graph.replay()
            return self.result
        if self.calls < self.warmup:
            self.calls += 1
            return self.fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.result = self.fn()
        self.graph = g
        g.replay()
        return self.result
