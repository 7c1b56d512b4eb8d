"""Edge cases of the path on the GPU against the oracle: minimal sizes, rays that miss, a camera inside the box, single-sample
rays, max_samples = 1 (H2), fully transparent and immediately opaque transfer functions, flat volumes (H4)."""
import numpy as np
import pytest
import torch

from helpers import GRAD_TOL, RGBA_TOL, oracle_backward_views, oracle_forward_views, rel_l2

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _run(vol, tf, cams, out_shape, jit=None, sr=1.0, M=2048, check_grad=True):
    from differender_b200 import VolumeRaycaster
    D, H, W = vol.shape[-3:]
    vr = VolumeRaycaster((W, D, H), out_shape, max_samples=M, tf_resolution=tf.shape[-1])
    v = vr.brick(vol.to(DEV).reshape(1, D, H, W).contiguous())
    tf_r4 = tf.to(DEV).t().contiguous()[None]
    c = cams.to(DEV).contiguous()
    j = None if jit is None else jit.to(DEV).contiguous()
    out, K, Tp = vr.march(v, tf_r4, c, sr, j)
    ref, Kr, nr = oracle_forward_views(vol, tf, cams, out_shape, jit, sampling_rate=sr, max_samples=M)
    assert np.array_equal(K.cpu().numpy(), Kr)
    assert np.abs(out.cpu().numpy() - ref).max() <= RGBA_TOL
    if check_grad:
        go = torch.randn(ref.shape, generator=torch.Generator().manual_seed(1))
        gv, gt = vr.march_backward(v, tf_r4, c, sr, j, go.to(DEV), out, K, Tp, True, True)
        gvr, gtr = oracle_backward_views(vol, tf, cams, go.numpy(), out_shape, jit, sampling_rate=sr, max_samples=M)
        assert torch.isfinite(gv).all() and torch.isfinite(gt).all()
        if np.abs(gvr).max() > 0:
            assert rel_l2(gv[0].cpu().numpy(), gvr) <= GRAD_TOL
        else:
            assert not gv.any()
        if np.abs(gtr).max() > 0:
            assert rel_l2(gt[0].cpu().numpy().T, gtr) <= GRAD_TOL
        else:
            assert not gt.any()
    return out, K, nr


def _noise_vol(shape, seed=0, lo=0.2, hi=0.8):
    g = torch.Generator().manual_seed(seed)
    return (lo + (hi - lo) * torch.rand((1,) + tuple(shape), generator=g))


def _tf(R, alpha=0.1, seed=0):
    g = torch.Generator().manual_seed(seed)
    t = torch.rand(4, R, generator=g)
    t[3] = alpha
    return t


CAM = torch.tensor([[1.2, 0.7, 2.2]])


def test_minimal_sizes():
    _run(_noise_vol((2, 2, 2)), _tf(2), CAM, (1, 1))
    _run(_noise_vol((2, 3, 5)), _tf(3), CAM, (3, 2))
    _run(_noise_vol((9, 8, 7)), _tf(5), CAM, (17, 9), jit=torch.rand(1, 9, 17, generator=torch.Generator().manual_seed(2)))


def test_all_rays_miss_and_far_camera():
    # from far away the box covers a handful of pixels; everything else must be exactly zero with K = 0
    out, K, n = _run(_noise_vol((16, 16, 16)), _tf(8), torch.tensor([[30.0, 9.0, 25.0]]), (24, 24))
    assert (n == 0).mean() > 0.9 and (out.permute(1, 0, 2, 3)[:, K == 0] == 0).all()


def test_camera_inside_the_box():
    # tmin < 0: the slab test still reports a hit (tmax >= 0); samples behind the camera are taken like the reference does
    _run(_noise_vol((24, 24, 24)), _tf(16, alpha=0.03), torch.tensor([[0.3, 0.2, 0.4]]), (20, 16))


def test_single_sample_rays_and_max_samples_one():
    vol = _noise_vol((12, 12, 12))
    out, K, n = _run(vol, _tf(8), CAM, (32, 32), sr=0.01)               # n == 1 for every hit ray (H3: sample at t0)
    assert set(np.unique(n)) <= {0, 1} and (n == 1).any()
    out, K, n = _run(vol, _tf(8), CAM, (16, 16), M=1)                   # H2: only the first sample is composited
    assert K.max().item() == 1 and n.max() > 1


def test_transparent_opaque_and_flat():
    vol = _noise_vol((16, 16, 16))
    out, K, n = _run(vol, _tf(8, alpha=0.0), CAM, (16, 16))             # nothing visible, every sample active and skipped
    assert (out == 0).all() and (K.cpu().numpy() == n).all()
    out, K, n = _run(vol, _tf(8, alpha=1.0), CAM, (16, 16))             # opaque at the first sample (H8: finite gradient)
    assert K.max().item() == 1
    flat = torch.full((1, 16, 16, 16), 0.5)                             # H4: zero gradient everywhere -> ambient-only shading
    out, K, n = _run(flat, _tf(8, alpha=0.05), CAM, (16, 16))
    assert torch.isfinite(out).all()
