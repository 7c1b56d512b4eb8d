"""world_size-2 gloo test of the multi-GPU host logic (view sharding + single all-reduce of [volume grad | TF grad]).
The CUDA renderer is replaced by a differentiable CPU stand-in with the Raycaster call signature; only the
distributed plumbing in differender_b200/distributed.py is under test here."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from differender_b200.distributed import DistributedRaycaster, shard_views


class _StubRaycaster(torch.nn.Module):
    """img[v] = f(volume, tf, cam_v): linear in volume and tf so the expected summed gradient is easy to state."""
    output_shape = (4, 3)

    def forward(self, volume, tf, look_from, jitter_tensor=None):
        s = volume.sum() * 0.5 + (tf * tf).sum()
        img = look_from.sum(dim=1).view(-1, 1, 1, 1) * s * torch.ones(look_from.shape[0], 4, 3, 4)
        if jitter_tensor is not None:
            img = img + jitter_tensor.unsqueeze(1)
        return img


class _SinkVolumeGrad(torch.autograd.Function):
    """volume -> volume.sum()*0.5, whose backward writes the volume gradient into the raycaster's grad sink when there is one
    (what VolumeRaycaster.march_backward does with the gather), so SyncGradients finds it already in place."""

    @staticmethod
    def forward(ctx, vr, volume):
        ctx.vr, ctx.shape = vr, volume.shape
        return volume.sum() * 0.5

    @staticmethod
    def backward(ctx, g):
        sink = ctx.vr.grad_sink.get("vol") if ctx.vr.grad_sink else None
        out = sink.view(ctx.shape) if sink is not None else torch.empty(ctx.shape)
        out.fill_(0.5 * float(g))
        ctx.vr.sunk = sink is not None
        return None, out


class _StubVr:
    volume_resolution = (5, 3, 4)          # (X, Y, Z) of a (1, 3, 4, 5) volume
    tf_resolution = 6
    defer_volume_gather = False
    grad_sink = None
    sunk = None


class _StubRaycasterWithSink(_StubRaycaster):
    def __init__(self):
        super().__init__()
        self.vr = _StubVr()

    def forward(self, volume, tf, look_from, jitter_tensor=None):
        s = _SinkVolumeGrad.apply(self.vr, volume) + (tf * tf).sum()
        img = look_from.sum(dim=1).view(-1, 1, 1, 1) * s * torch.ones(look_from.shape[0], 4, 3, 4)
        if jitter_tensor is not None:
            img = img + jitter_tensor.unsqueeze(1)
        return img


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_views, ret, sink=False):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    vol = torch.rand(1, 3, 4, 5, requires_grad=True)
    tf = torch.rand(4, 6, requires_grad=True)
    cams = torch.arange(n_views * 3, dtype=torch.float32).view(n_views, 3) / 10
    jit = torch.rand(n_views, 3, 4)
    drc = DistributedRaycaster(_StubRaycasterWithSink() if sink else _StubRaycaster())
    # the jitter may be given for every view or already sharded (this rank's views, in shard order)
    idx0 = shard_views(n_views, rank, world)
    img, idx = drc(vol, tf, cams, jit[idx0] if (sink and idx0) else jit)
    assert idx == idx0 and img.shape[0] == len(idx)
    img.sum().backward()
    if sink and idx:
        assert drc.raycaster.vr.sunk is True                   # the backward found the flat buffer's slot and wrote there
        assert vol.grad.shape == vol.shape
    # single-process reference: all views on one rank
    v2 = vol.detach().clone().requires_grad_(True); t2 = tf.detach().clone().requires_grad_(True)
    _StubRaycaster()(v2, t2, cams, jit).sum().backward()
    ok = torch.allclose(vol.grad, v2.grad, rtol=1e-5) and torch.allclose(tf.grad, t2.grad, rtol=1e-5)
    ret[rank] = bool(ok)
    dist.barrier()
    dist.destroy_process_group()


def _run(world, n_views, sink=False):
    ctx = mp.get_context("spawn")
    with ctx.Manager() as m:
        ret = m.dict()
        port = _free_port()
        procs = [ctx.Process(target=_worker, args=(r, world, port, n_views, ret, sink)) for r in range(world)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(120)
            assert p.exitcode == 0
        assert all(ret.get(r) for r in range(world)), dict(ret)


def test_two_ranks_sum_gradients_over_all_views():
    _run(2, 5)


def test_rank_without_views_still_joins_the_collective():
    _run(2, 1)


def test_backward_writes_into_the_flat_buffer_without_a_copy():
    _run(2, 5, sink=True)


def test_views_are_dealt_round_robin():
    assert shard_views(5, 0, 2) == [0, 2, 4] and shard_views(5, 1, 2) == [1, 3]
    assert shard_views(1, 1, 2) == [] and shard_views(16, 3, 8) == [3, 11]
    got = sorted(v for r in range(8) for v in shard_views(61, r, 8))
    assert got == list(range(61))
