"""View-sharded multi-GPU rendering.

The reference is single-GPU and walks a batch of views in a Python loop (volume_raycaster.py:418-426, 450-464).  Views are
independent, so the batch shards across ranks with the volume and the transfer function replicated; the only exchange is
the gradient sum: ONE all-reduce (NCCL over NVLink on GPUs, gloo in the CPU tests) of a flat buffer
[volume gradient | TF gradient].  One process per GPU (torchrun); nothing here spawns processes.

The flat buffer is allocated once and the gather of the cell-major volume gradient writes into it directly, so the volume
gradient (64 MiB - 4 GiB) is neither concatenated nor copied before the collective; the 2 KiB TF gradient is copied into its slot.
"""
import torch
import torch.distributed as dist

__all__ = ["shard_views", "SyncGradients", "DistributedRaycaster", "GradBuffer"]


def shard_views(n_views, rank, world_size):
    """Indices of the views rendered by `rank`: round-robin, view v -> rank v mod world_size.  Neighbouring
    cameras of an orbit cost about the same, so dealing them out one by one spreads the expensive poses over all ranks
    (contiguous blocks left the ranks 1.3 % apart in samples at 8 GPUs); sizes differ by at most one."""
    return list(range(rank, n_views, world_size))


class GradBuffer:
    """A flat fp32 gradient buffer and its all-reduce (SUM) over the ranks of one box.

    Measured on 8 x B200 (profiles/r02_allreduce_n8.txt): NCCL's all-reduce reaches 393 / 695 / 831 GB/s bus bandwidth at
    64 MiB / 512 MiB / 4 GiB; the NVLS multimem all-reduce on a symmetric-memory buffer (torch.distributed._symmetric_memory:
    the reduction happens in the NVSwitch) reaches 677 GB/s at 64 MiB -- 0.17 instead of 0.30 ms for the C3 gradient -- and ties
    at 512 MiB.  So buffers up to MULTIMEM_MAX_BYTES on at least MULTIMEM_MIN_RANKS GPUs are allocated as symmetric memory and
    reduced with multimem; everything else (and any box where symmetric memory is not available: gloo, older drivers) uses the
    process group's all_reduce.  Allocation is collective: every rank must create its buffers in the same order."""
    MULTIMEM_MAX_BYTES = 256 << 20
    MULTIMEM_MIN_RANKS = 4

    def __init__(self, numel, device, group=None):
        self.group, self.how = group, "all_reduce"
        self.buf = None
        world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        dev = torch.device(device)
        if dev.type == "cuda" and world >= self.MULTIMEM_MIN_RANKS and numel * 4 <= self.MULTIMEM_MAX_BYTES and dist.get_backend(group) == "nccl":
            try:
                import torch.distributed._symmetric_memory as symm
                name = (group or dist.group.WORLD).group_name
                t = symm.empty(numel, dtype=torch.float32, device=dev)
                symm.rendezvous(t, name)
                if hasattr(torch.ops.symm_mem, "multimem_all_reduce_"):
                    self.buf, self.how, self._name = t, "multimem_all_reduce (NVLS, symmetric memory)", name
            except Exception:            # symmetric memory is an optimisation: any failure to set it up means the NCCL path
                self.buf = None
        if self.buf is None:
            self.buf = torch.empty(numel, dtype=torch.float32, device=dev)

    def all_reduce(self, part=None):
        """Sums `part` (a prefix slice of the buffer; default: all of it) over the ranks, in place."""
        if part is None and self.how != "all_reduce":
            torch.ops.symm_mem.multimem_all_reduce_(self.buf, "sum", self._name)
        else:
            dist.all_reduce(self.buf if part is None else part, op=dist.ReduceOp.SUM, group=self.group)


class _FlatGrads:
    """One preallocated fp32 buffer [g_0 | g_1 | ...] for the gradients of a fixed list of parameter shapes."""

    def __init__(self, group=None):
        self.gb, self.sizes, self.group = None, None, group

    @property
    def buf(self):
        return self.gb.buf

    def views(self, tensors):
        sizes = tuple(int(t.numel()) for t in tensors)
        dev = tensors[0].device
        if self.gb is None or self.sizes != sizes or self.gb.buf.device != dev:
            self.gb, self.sizes = GradBuffer(sum(sizes), dev, self.group), sizes
        out, off = [], 0
        for n in sizes:
            out.append(self.gb.buf[off:off + n])
            off += n
        return out


class SyncGradients(torch.autograd.Function):
    """Identity in the forward; in the backward the gradients of all inputs are summed over the ranks with a single
    all-reduce of one preallocated flat fp32 buffer.  A gradient that already lives at its place in that buffer (the
    raycaster's backward writes there, see DistributedRaycaster) is not copied; anything else is copied in once.
    Replicated parameters (volume, TF) pass through this before the per-rank render so that every rank ends up with the
    gradient of the whole view batch."""

    @staticmethod
    def forward(ctx, group, flat, *tensors):
        ctx.group, ctx.flat = group, flat
        ctx.set_materialize_grads(False)     # an input without a gradient (e.g. a volume whose gradient FusedVolumeSGD keeps cell-major) stays None
        return tuple(t.view_as(t) for t in tensors)

    @staticmethod
    def backward(ctx, *grads):
        if all(g is None for g in grads):
            return (None, None) + tuple(grads)
        if not (dist.is_available() and dist.is_initialized() and dist.get_world_size(ctx.group) > 1):
            return (None, None) + tuple(grads)
        flat = ctx.flat if ctx.flat is not None else _FlatGrads(ctx.group)
        present = [g for g in grads if g is not None]
        slots = flat.views(present)
        for g, s in zip(present, slots):
            if not (g.dtype == torch.float32 and g.is_contiguous() and g.data_ptr() == s.data_ptr()):
                s.copy_(g.reshape(-1))                      # not written in place by the backward kernels: one copy, no concatenation
        flat.gb.all_reduce()
        out, it = [], iter(slots)
        for g in grads:
            out.append(None if g is None else next(it).view(g.shape).to(g.dtype))
        return (None, None) + tuple(out)


class DistributedRaycaster(torch.nn.Module):
    """Wraps a `Raycaster`: each rank renders its shard of the cameras; volume/TF gradients are summed over ranks.

    forward(volume, tf, look_from_all[, jitter]) -> this rank's images ([n_local, 4, H, W]) and the view indices.
    `jitter` holds either every view's jitter ([n_views, H, W]) or only this rank's ([n_local, H, W], in shard order).
    """

    def __init__(self, raycaster, group=None):
        super().__init__()
        self.raycaster = raycaster
        self.group = group
        self._flat = _FlatGrads(group)

    def _rank_world(self):
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(self.group), dist.get_world_size(self.group)
        return 0, 1

    def forward(self, volume, tf, look_from_all, jitter=None):
        return self._render(volume, tf, look_from_all, jitter, None)

    def mse_loss(self, volume, tf, look_from_all, target, jitter=None):
        """Fused render + MSE of this rank's views against `target` ([n_local, 4, H, W]): returns (sum of squared errors over
        this rank's views / (n_views_total * 4 * H * W), images, view indices) -- the per-rank terms add up to the batch's mean-squared
        error, so after loss.backward() every rank holds the gradient of the whole batch's MSE (all-reduced inside the backward)."""
        return self._render(volume, tf, look_from_all, jitter, target)

    def _render(self, volume, tf, look_from_all, jitter, target):
        rank, world = self._rank_world()
        idx = shard_views(look_from_all.shape[0], rank, world)
        need = [t for t in (volume, tf) if t.requires_grad]
        vr = getattr(self.raycaster, "vr", None)
        if vr is not None:
            vr.grad_sink = None
            if world > 1 and volume.requires_grad and volume.ndim == 4 and not getattr(vr, "defer_volume_gather", False):
                # a shared volume: the backward gathers its gradient straight into the flat buffer that is all-reduced
                # (the TF gradient, 2 KiB, comes back through a permute and is copied into its slot)
                X, Y, Z = vr.volume_resolution
                vr.grad_sink = {"vol": self._flat.views(need)[0].view(1, Y, Z, X)}
        volume, tf = SyncGradients.apply(self.group, self._flat, volume, tf)
        if not idx:
            h, w = self.raycaster.output_shape[1], self.raycaster.output_shape[0]
            # keep the graph connected so the collective in the backward still runs on this rank
            empty = volume.new_zeros((0, 4, h, w), dtype=torch.float32) + 0.0 * (volume.sum() + tf.sum()).float()
            return (empty.sum(), empty, idx) if target is not None else (empty, idx)
        sel = torch.arange(rank, look_from_all.shape[0], world, device=look_from_all.device)    # built on the device: no synchronising host copy
        jit = jitter
        if jitter is not None and jitter.shape[0] != len(idx):
            jit = jitter.index_select(0, sel.to(jitter.device))
        cams = look_from_all.index_select(0, sel)
        if target is not None:
            loss, img = self.raycaster.mse_loss(volume, tf, cams, target, jit)
            return loss * (len(idx) / look_from_all.shape[0]), img, idx      # mean over this rank's views -> this rank's share of the batch mean
        img = self.raycaster(volume, tf, cams, jit)
        return img, idx
