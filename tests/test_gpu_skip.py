"""Exact empty-space skipping (dr_build_skip_grid / dr_forward_ex): the forward with the skip grid must reproduce the forward
without it bit for bit -- image, active-sample counts K and Tprev -- in every layout, dtype and march variant, and the grid the
device builds must equal the one the host restatement of the same functions builds."""
import numpy as np
import pytest
import torch

import hostsim_lib as hs
from helpers import case_inputs, rel_l2

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _march(vol, tf, cams, jit, out_shape, layout, skip, dtype=torch.float32, sr=1.0, nondiff=False, M=4096, batched_tf=False):
    from differender_b200 import VolumeRaycaster
    D, H, W = vol.shape[-3:]
    vr = VolumeRaycaster((W, D, H), out_shape, max_samples=M, tf_resolution=tf.shape[-1], layout=layout, skip_empty=skip)
    bricked = vr.brick(vol.to(DEV, dtype).reshape(1, D, H, W).contiguous())
    tf_r4 = tf.to(DEV).t().contiguous()[None]
    if batched_tf:                                           # one TF per view: view 1 gets a shifted copy
        tf_r4 = torch.cat([tf_r4, torch.roll(tf_r4, 7, dims=1)]).contiguous()
    return vr.march(bricked, tf_r4, cams.to(DEV).contiguous(), sr, None if jit is None else jit.to(DEV).contiguous(), nondiff=nondiff)


@pytest.mark.parametrize("layout", ["linear", "brick8", "cell8"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
def test_skip_is_bit_identical(layout, dtype):
    vol, tf, cams, jit = case_inputs((72, 64, 80), (96, 64), 128, seed=31, tf_name="tf1", views=2)
    a = _march(vol, tf, cams, jit, (96, 64), layout, False, dtype)
    b = _march(vol, tf, cams, jit, (96, 64), layout, True, dtype)
    for x, y in zip(a, b):
        assert torch.equal(x, y)
    assert a[1].sum().item() > 0


@pytest.mark.parametrize("kw", [dict(sr=0.7), dict(sr=4.0, nondiff=True), dict(batched_tf=True), dict(tf_name="tf5"), dict(tf_name="gray")])
def test_skip_variants_are_bit_identical(kw):
    kw = dict(kw)
    tf_name = kw.pop("tf_name", "tf1")
    vol, tf, cams, jit = case_inputs((64, 64, 64), (80, 56), 128, seed=17, tf_name=tf_name, views=2)
    a = _march(vol, tf, cams, jit, (80, 56), "cell8", False, **kw)
    b = _march(vol, tf, cams, jit, (80, 56), "cell8", True, **kw)
    for x, y in zip(a, b):
        assert (x is None and y is None) or torch.equal(x, y)


def test_long_axis_and_gradients_unchanged():
    # an axis of 1100 voxels (TAPS_TWO) through the autograd API: skipping only touches the forward, gradients must not move
    from differender_b200 import Raycaster
    vol, tf, cams, jit = case_inputs((9, 1100, 9), (32, 24), 64, seed=9, tf_name="tf1", views=1)
    res = []
    for skip in (False, True):
        rc = Raycaster((9, 1100, 9), (32, 24), 64, max_samples=4096, skip_empty=skip)
        v = vol.to(DEV).requires_grad_(True); t = tf.to(DEV).requires_grad_(True)
        img = rc(v, t, cams[0].to(DEV), jit[0].to(DEV))
        (img * torch.linspace(0, 1, img.numel(), device=DEV).reshape(img.shape)).sum().backward()
        res.append((img.detach(), v.grad.clone(), t.grad.clone()))
    assert torch.equal(res[0][0], res[1][0])
    for x, y in zip(res[0][1:], res[1][1:]):               # float atomics: the summation order differs from run to run
        assert rel_l2(y.cpu().numpy(), x.cpu().numpy()) <= 1e-5


def test_device_grid_equals_host_grid():
    import ctypes
    from differender_b200 import VolumeRaycaster, _lib
    vol, tf, _, _ = case_inputs((40, 56, 72), (8, 8), 64, seed=4, tf_name="tf3", jitter=False)
    vr = VolumeRaycaster((72, 40, 56), (8, 8), max_samples=64, tf_resolution=64)
    lin = vol.to(DEV).reshape(1, 40, 56, 72).contiguous()
    tf_r4 = tf.to(DEV).t().contiguous()[None]
    d = vr.desc(1, 1, 1, _lib.VOX_F32, 0, 1.0)
    grid = vr.skip_grid(d, lin, tf_r4)
    ref = hs.skip_grid(vol.numpy(), tf.numpy(), (8, 8), max_samples=64)
    g = grid.cpu().numpy()
    assert g.size == 16 + ref.size and ref.size in (5 * 7 * 9, 10 * 14 * 18)          # 8^3- or 4^3-cell macro-cells of a 40 x 56 x 72 volume
    assert np.array_equal(g[16:].reshape(ref.shape), ref) and 0 < ref.mean() < 1
    assert int(g[:4].view(np.uint32)[0]) == int(ref.sum())          # header: number of empty macro-cells


@pytest.mark.parametrize("layout,dtype,shape", [("cell8", torch.float32, (72, 64, 80)), ("linear", torch.float32, (72, 64, 80)),
                                                ("cell8", torch.float16, (9, 1100, 9)), ("brick8", torch.float32, (40, 56, 72))])
def test_volume_only_backward_skips_exactly(layout, dtype, shape):
    # dr_backward_ex with the forward's grid: the volume-only backward jumps over empty macro-cells; the gradient must equal the
    # full march's (to the rounding of the atomics' order) and the oracle's
    from differender_b200 import VolumeRaycaster
    from helpers import GRAD_TOL, oracle_backward_views, oracle_forward_views
    out_shape = (64, 48)
    vol, tf, cams, jit = case_inputs(shape, out_shape, 128, seed=33, tf_name="tf1", views=2)
    D, H, W = shape
    vr = VolumeRaycaster((W, D, H), out_shape, max_samples=4096, tf_resolution=128, layout=layout)
    b = vr.brick(vol.to(DEV, dtype).reshape(1, D, H, W).contiguous())
    tf_r4 = tf.to(DEV).t().contiguous()[None]
    c, j = cams.to(DEV).contiguous(), jit.to(DEV).contiguous()
    out, K, Tp = vr.march(b, tf_r4, c, 1.0, j)
    grid = vr.last_skip_grid
    assert grid is not None and int(grid[:4].view(torch.int32)[0]) > 0
    go = torch.randn(out.shape, generator=torch.Generator().manual_seed(4)).to(DEV)
    gv0, _ = vr.march_backward(b, tf_r4, c, 1.0, j, go, out, K, Tp, True, False)
    gv1, gt1 = vr.march_backward(b, tf_r4, c, 1.0, j, go, out, K, Tp, True, False, skip_grid=grid)
    assert gt1 is None and gv0.abs().max().item() > 0
    assert rel_l2(gv1.cpu().numpy(), gv0.cpu().numpy()) <= 1e-5
    # with a TF gradient the grid is ignored
    gv2, gt2 = vr.march_backward(b, tf_r4, c, 1.0, j, go, out, K, Tp, True, True, skip_grid=grid)
    gv3, gt3 = vr.march_backward(b, tf_r4, c, 1.0, j, go, out, K, Tp, True, True)
    assert rel_l2(gv2.cpu().numpy(), gv3.cpu().numpy()) <= 1e-5 and rel_l2(gt2.cpu().numpy(), gt3.cpu().numpy()) <= 1e-5
    if dtype == torch.float32 and max(shape) < 1000:
        ref, _, _ = oracle_forward_views(vol, tf, cams, out_shape, jit, max_samples=4096)
        gvr, _ = oracle_backward_views(vol, tf, cams, go.cpu().numpy(), out_shape, jit, max_samples=4096, want_tf=False)
        assert rel_l2(gv1[0].cpu().numpy(), gvr) <= GRAD_TOL


def test_volume_only_autograd_uses_the_grid_and_matches():
    from differender_b200 import Raycaster
    vol, tf, cams, jit = case_inputs((64, 64, 64), (48, 40), 128, seed=35, tf_name="tf1", views=2)
    res = []
    for skip in (False, True):
        rc = Raycaster((64, 64, 64), (48, 40), 128, max_samples=2048, skip_empty=skip)
        v = vol.to(DEV).requires_grad_(True)
        img = rc(v, tf.to(DEV), cams.to(DEV), jit.to(DEV))                    # the TF does not require grad: volume-only backward
        (img * torch.linspace(0, 1, img.numel(), device=DEV).reshape(img.shape)).sum().backward()
        res.append((img.detach(), v.grad.clone()))
    assert torch.equal(res[0][0], res[1][0])
    assert rel_l2(res[1][1].cpu().numpy(), res[0][1].cpu().numpy()) <= 1e-5
