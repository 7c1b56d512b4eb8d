"""Caller-side optimiser step of the reference's TF-optimisation demo, fused into one elementwise kernel
(reference examples/taichi_volume_raycaster.py:375-381 `apply_grad` and :596-602 the loop)."""
import ctypes

import torch

from . import _lib

__all__ = ["MomentumSGD", "FusedVolumeSGD"]


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


class FusedVolumeSGD:
    """Volume-optimisation step with the gather fused in (dr_gather_step): ONE kernel reads the
    cell-major gradient the backward scattered, applies nan_to_num, gradient clipping, momentum and the projection
    `vol.clamp_(lo, hi)` (reference examples/test_opt_tf.py:86-88; update rule of examples/taichi_volume_raycaster.py:375-381),
    writes the parameter and the momentum in place AND refreshes the cell-major copy of the volume the next forward reads --
    so a step launches no gather_grad / momentum_step / expand_cells kernels of its own.

        rc = Raycaster(...); vol = torch.nn.Parameter-like fp32 (1, D, H, W) CUDA tensor with requires_grad
        opt = FusedVolumeSGD(rc, vol, lr=..., hi=1.0)
        loss = f(rc(vol, tf, cams)); loss.backward(); opt.step()          # vol.grad stays None: the gradient never leaves cell-major form

    With torch.distributed initialised (`group` given or world size > 1) the gradient is gathered into a flat buffer, all-reduced
    (one NCCL call) and the step runs from the reduced linear gradient (dr_gather_step's `grad_linear` input)."""

    def __init__(self, raycaster, param, lr=0.1, momentum=0.9, max_grad=0.1, lr_decay=0.99, lo=0.0, hi=1.0, group=None):
        vr = getattr(raycaster, "vr", raycaster)
        if not param.is_cuda or param.dtype != torch.float32 or not param.is_contiguous():
            raise RuntimeError("FusedVolumeSGD needs a contiguous fp32 CUDA volume (no CPU fallback)")
        X, Y, Z = vr.volume_resolution
        if param.numel() != X * Y * Z:
            raise ValueError(f"volume has {param.numel()} voxels, the raycaster was built for {X * Y * Z}")
        self.vr, self.param, self.group = vr, param, group
        self.lr, self.gamma, self.max_grad, self.lr_decay, self.lo, self.hi = lr, momentum, max_grad, lr_decay, lo, hi
        self.state = torch.zeros_like(param)
        self._flat = None
        vr.defer_volume_gather = True

    def _world(self):
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_world_size(self.group)
        return 1

    @torch.no_grad()
    def step(self, grad_out=None):
        """Applies the pending cell-major gradient (accumulated by every backward since the last step).  `grad_out`: optional
        fp32 tensor of the volume's shape that receives the gathered (and, multi-GPU, reduced) gradient."""
        vr, p = self.vr, self.param
        cells = vr.pending_grad_cells
        if cells is None:
            raise RuntimeError("FusedVolumeSGD.step: no pending volume gradient (call loss.backward() first)")
        X, Y, Z = vr.volume_resolution
        lin = p.view(1, Y, Z, X)                                  # the (1, D, H, W) tensor is [Y][Z][X] in the library's axis names
        cached = vr._cached_copy(lin)                             # the cell-major copy the forward made of this very tensor, if any
        if cached is not None and cached.ndim != 3:
            cached = None                                         # a bricked copy is rebuilt by the next forward instead
        vox = _lib.VOX_F16 if (cached is not None and cached.dtype == torch.float16) else _lib.VOX_F32
        d = vr.desc(1, 1, 1, vox, 0, 1.0)
        with torch.cuda.device(p.device):
            st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
            g_cells, g_lin = cells, None
            if self._world() > 1:
                from .distributed import GradBuffer
                if self._flat is None:
                    self._flat = GradBuffer(p.numel(), p.device, self.group)       # NVLS multimem up to 256 MiB, NCCL above
                vr.gather(cells, out=self._flat.buf.view(1, Y, Z, X))
                self._flat.all_reduce()
                g_cells, g_lin = None, self._flat.buf
            if grad_out is not None and (grad_out.dtype != torch.float32 or not grad_out.is_contiguous() or grad_out.numel() != p.numel()):
                raise ValueError("grad_out must be a contiguous fp32 tensor of the volume's size")
            _lib.check(_lib.load().dr_gather_step(ctypes.byref(d), _lib.ptr(g_cells), _lib.ptr(g_lin), _lib.ptr(p), _lib.ptr(self.state),
                                                  _lib.ptr(cached), _lib.ptr(grad_out), self.lr, self.gamma, self.max_grad, self.lo, self.hi, st),
                       "dr_gather_step")
            cells.zero_()
        torch.autograd.graph.increment_version(p)                 # written behind PyTorch's back
        if cached is not None:                                    # the refreshed copy now belongs to the parameter's new version
            vr._copy_cache = (vr._src_key(lin), lin.untyped_storage(), cached)
            cached.dr_source = lin
        self.lr *= self.lr_decay
