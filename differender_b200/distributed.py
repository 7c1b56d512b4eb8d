"""View-sharded multi-GPU rendering.

The reference is single-GPU and walks a batch of views in a Python loop (volume_raycaster.py:418-426, 450-464).  Views are
independent, so the batch shards across ranks with the volume and the transfer function replicated; the only exchange is
the gradient sum: ONE all-reduce (NCCL over NVLink on GPUs, gloo in the CPU tests) of a flat buffer
[volume gradient | TF gradient].  One process per GPU (torchrun); nothing here spawns processes.
"""
import torch
import torch.distributed as dist

__all__ = ["shard_views", "SyncGradients", "DistributedRaycaster"]


def shard_views(n_views, rank, world_size):
    """Indices of the views rendered by `rank`: contiguous blocks, sizes differ by at most one."""
    base, rem = divmod(n_views, world_size)
    start = rank * base + min(rank, rem)
    return list(range(start, start + base + (1 if rank < rem else 0)))


class SyncGradients(torch.autograd.Function):
    """Identity in the forward; in the backward the gradients of all inputs are packed into one flat fp32 buffer and
    all-reduced (SUM) with a single collective, then unpacked.  Replicated parameters (volume, TF) pass through this
    before the per-rank render so that every rank ends up with the gradient of the whole view batch."""

    @staticmethod
    def forward(ctx, group, *tensors):
        ctx.group = group
        return tuple(t.view_as(t) for t in tensors)

    @staticmethod
    def backward(ctx, *grads):
        present = [g for g in grads if g is not None]
        if not present:
            return (None,) + tuple(grads)
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(ctx.group) > 1:
            flat = torch.cat([g.reshape(-1).float() for g in present])
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=ctx.group)
            out, off = [], 0
            for g in grads:
                if g is None:
                    out.append(None)
                    continue
                n = g.numel()
                out.append(flat[off:off + n].view(g.shape).to(g.dtype))
                off += n
            return (None,) + tuple(out)
        return (None,) + tuple(grads)


class DistributedRaycaster(torch.nn.Module):
    """Wraps a `Raycaster`: each rank renders its shard of the cameras; volume/TF gradients are summed over ranks.

    forward(volume, tf, look_from_all[, jitter_all]) -> this rank's images ([n_local, 4, H, W]) and the view indices.
    """

    def __init__(self, raycaster, group=None):
        super().__init__()
        self.raycaster = raycaster
        self.group = group

    def _rank_world(self):
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(self.group), dist.get_world_size(self.group)
        return 0, 1

    def forward(self, volume, tf, look_from_all, jitter_all=None):
        rank, world = self._rank_world()
        idx = shard_views(look_from_all.shape[0], rank, world)
        volume, tf = SyncGradients.apply(self.group, volume, tf)
        if not idx:
            h, w = self.raycaster.output_shape[1], self.raycaster.output_shape[0]
            # keep the graph connected so the collective in the backward still runs on this rank
            empty = volume.new_zeros((0, 4, h, w), dtype=torch.float32) + 0.0 * (volume.sum() + tf.sum()).float()
            return empty, idx
        sel = torch.as_tensor(idx, device=look_from_all.device)
        jit = None if jitter_all is None else jitter_all.index_select(0, sel.to(jitter_all.device))
        img = self.raycaster(volume, tf, look_from_all.index_select(0, sel), jit)
        return img, idx
