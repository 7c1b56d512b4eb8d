"""Caller-side optimiser step of the reference's TF-optimisation demo, fused into one elementwise kernel
(reference examples/taichi_volume_raycaster.py:375-381 `apply_grad` and :596-602 the loop)."""
import ctypes

import torch

from . import _lib

__all__ = ["MomentumSGD"]


class MomentumSGD:
    """m = gamma*m + lr*clamp(grad, +-max_grad);  p = clamp(p - m, lo, hi);  lr *= lr_decay after every step.

    Defaults are the demo's (--lr 0.1 --mom 0.9 --clip-grads 0.1 --lr-decay 0.99, :500-516) with the TF projection
    `max(tf, 0)`; pass hi=1.0 for the volume projection `vol.clamp_(0, 1)` of examples/test_opt_tf.py:86-88."""

    def __init__(self, param, lr=0.1, momentum=0.9, max_grad=0.1, lr_decay=0.99, lo=0.0, hi=float("inf")):
        if not param.is_cuda or param.dtype != torch.float32 or not param.is_contiguous():
            raise RuntimeError("MomentumSGD needs a contiguous fp32 CUDA tensor (no CPU fallback)")
        self.param, self.lr, self.gamma, self.max_grad, self.lr_decay, self.lo, self.hi = param, lr, momentum, max_grad, lr_decay, lo, hi
        self.state = torch.zeros_like(param)

    @torch.no_grad()
    def step(self, grad=None):
        g = self.param.grad if grad is None else grad
        if g is None:
            raise RuntimeError("no gradient to apply")
        if not g.is_cuda or g.device != self.param.device or g.numel() != self.param.numel():
            raise RuntimeError(f"MomentumSGD.step: gradient must be a CUDA tensor on {self.param.device} with {self.param.numel()} "
                               f"elements (got {tuple(g.shape)} on {g.device}); the kernel reads one gradient element per parameter element")
        g = g.float().contiguous()
        with torch.cuda.device(self.param.device):
            st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
            _lib.check(_lib.load().dr_momentum_step(_lib.ptr(self.param), _lib.ptr(g), _lib.ptr(self.state), self.param.numel(),
                                                    self.lr, self.gamma, self.max_grad, self.lo, self.hi, st), "dr_momentum_step")
        # the kernel wrote the parameter through its raw pointer: tell PyTorch (autograd's in-place checks, and the caches of
        # VolumeRaycaster, which are keyed on the version counter) that the tensor changed
        torch.autograd.graph.increment_version(self.param)
        self.lr *= self.lr_decay
