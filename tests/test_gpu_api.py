"""The public PyTorch API (Raycaster / RaycastFunction, mirroring reference differender/volume_raycaster.py:392-574) on the
GPU: shapes, batching semantics, autograd to volume and TF, checked against the oracle."""
import numpy as np
import pytest
import torch

from helpers import GRAD_TOL, RGBA_TOL, case_inputs, oracle_backward_views, oracle_forward_views, rel_l2

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _rc(vol, out_shape, R, **kw):
    from differender_b200 import Raycaster
    kw.setdefault("max_samples", 2048)
    return Raycaster(tuple(vol.shape[-3:]), out_shape, R, **kw)


def test_shared_volume_batched_cameras_autograd_matches_oracle():
    out_shape = (56, 40)
    vol, tf, cams, jit = case_inputs((40, 36, 44), out_shape, 64, seed=21, tf_name="tf1", views=3)
    rc = _rc(vol, out_shape, 64)
    v = vol.to(DEV).requires_grad_(True); t = tf.to(DEV).requires_grad_(True)
    img = rc(v, t, cams.to(DEV), jit.to(DEV))
    assert img.shape == (3, 4, 40, 56) and img.dtype == torch.float32 and img.is_contiguous()
    ref, Kr, _ = oracle_forward_views(vol, tf, cams, out_shape, jit, max_samples=2048)
    same = rc.vr.last_K.cpu().numpy() == Kr
    assert (~same).mean() <= 1e-4
    assert np.moveaxis(np.abs(img.detach().cpu().numpy() - ref), 1, 0)[:, same].max() <= RGBA_TOL
    go = torch.randn(img.shape, generator=torch.Generator().manual_seed(5))
    (img * go.to(DEV)).sum().backward()
    gv, gt = oracle_backward_views(vol, tf, cams, go.numpy(), out_shape, jit, max_samples=2048)
    assert v.grad.shape == vol.shape and t.grad.shape == tf.shape              # ONE summed gradient for the shared inputs
    assert rel_l2(v.grad[0].cpu().numpy(), gv) <= GRAD_TOL and rel_l2(t.grad.cpu().numpy(), gt) <= GRAD_TOL


def test_single_view_and_per_view_volumes_and_tfs():
    out_shape = (40, 32)
    vol, tf, cams, jit = case_inputs((32, 32, 32), out_shape, 32, seed=22, views=2)
    rc = _rc(vol, out_shape, 32)
    # un-batched: (1,D,H,W), (4,R), (3,) -> (4,H,W)   (reference :525-548)
    img = rc(vol.to(DEV), tf.to(DEV), cams[0].to(DEV), jit[0].to(DEV))
    assert img.shape == (4, 32, 40)
    ref, Kr, _ = oracle_forward_views(vol, tf, cams[:1], out_shape, jit[:1], max_samples=2048)
    assert np.abs(img.cpu().numpy() - ref[0]).max() <= RGBA_TOL
    # fully batched: each item has its own volume and TF; gradients come back per item (reference :447-449)
    vols = torch.stack([vol, torch.flip(vol, (2,))]).contiguous()               # (2,1,D,H,W)
    tfs = torch.stack([tf, tf.roll(3, dims=1)]).contiguous()                    # (2,4,R)
    V = vols.to(DEV).requires_grad_(True); T = tfs.to(DEV).requires_grad_(True)
    imgs = rc(V, T, cams.to(DEV), jit.to(DEV))
    go = torch.randn(imgs.shape, generator=torch.Generator().manual_seed(6))
    (imgs * go.to(DEV)).sum().backward()
    assert V.grad.shape == vols.shape and T.grad.shape == tfs.shape
    for i in range(2):
        r, Kr, _ = oracle_forward_views(vols[i], tfs[i], cams[i:i + 1], out_shape, jit[i:i + 1], max_samples=2048)
        assert np.abs(imgs[i].detach().cpu().numpy() - r[0]).max() <= RGBA_TOL
        gv, gt = oracle_backward_views(vols[i], tfs[i], cams[i:i + 1], go[i:i + 1].numpy(), out_shape, jit[i:i + 1], max_samples=2048)
        assert rel_l2(V.grad[i, 0].cpu().numpy(), gv) <= GRAD_TOL and rel_l2(T.grad[i].cpu().numpy(), gt) <= GRAD_TOL


def test_raycast_function_reference_calling_convention():
    # RaycastFunction.apply(vr, volume(X,Y,Z), tf(R,4), look_from, sampling_rate, (batched, bs), jitter) -> (W,H,4)   (:392-438)
    from differender_b200 import RaycastFunction
    out_shape = (40, 24)
    vol, tf, cams, _ = case_inputs((32, 32, 32), out_shape, 16, seed=23, views=1, jitter=False)
    rc = _rc(vol, out_shape, 16, jitter=False)
    _, _, vol_in, tf_in, lf_in = rc._determine_batch(vol.to(DEV), tf.to(DEV), cams[0].to(DEV))
    raw = RaycastFunction.apply(rc.vr, vol_in, tf_in, lf_in, 1.0, (False, 0), False)
    assert raw.shape == (40, 24, 4)
    img = torch.flip(raw, (1,)).permute(2, 1, 0).contiguous()                   # the reference's own post-processing (:543-548)
    assert torch.equal(img, rc(vol.to(DEV), tf.to(DEV), cams[0].to(DEV)))
    ref, _, _ = oracle_forward_views(vol, tf, cams, out_shape, None, max_samples=2048)
    assert np.abs(img.cpu().numpy() - ref[0]).max() <= RGBA_TOL


def test_nondiff_default_rate_and_no_grad():
    out_shape = (48, 48)
    vol, tf, cams, _ = case_inputs((48, 48, 48), out_shape, 128, seed=24, tf_name="tf1", views=2, jitter=False)
    rc = _rc(vol, out_shape, 128, sampling_rate=0.5)
    img = rc.raycast_nondiff(vol.to(DEV).requires_grad_(True), tf.to(DEV), cams.to(DEV))
    assert img.shape == (2, 4, 48, 48) and not img.requires_grad
    ref, _, _ = oracle_forward_views(vol, tf, cams, out_shape, None, sampling_rate=2.0, nondiff=True)     # 4 x 0.5 (:493)
    assert np.abs(img.cpu().numpy() - ref).max() <= RGBA_TOL


def test_internal_jitter_autocast_and_two_forwards_before_backward():
    out_shape = (32, 32)
    vol, tf, cams, jit = case_inputs((32, 32, 32), out_shape, 32, seed=25, views=1)
    rc = _rc(vol, out_shape, 32, jitter=True)
    v, t, c = vol.to(DEV), tf.to(DEV), cams[0].to(DEV)
    a, b = rc(v, t, c), rc(v, t, c)
    assert not torch.equal(a, b)                                                 # fresh uniform jitter per call (ti.random, :255)
    j = jit[0].to(DEV)
    assert torch.equal(rc(v, t, c, j), rc(v, t, c, j))                           # supplied jitter: deterministic
    with torch.autocast("cuda", dtype=torch.float16):
        h = rc(v, t, c, j)
    assert h.dtype == torch.float32 and torch.equal(h, rc(v, t, c, j))           # AMP-safe (:394, :441)
    # H10: state lives in ctx, so a second forward does not corrupt the first one's backward
    t1 = t.clone().requires_grad_(True)
    img1 = rc(v, t1, c, j)
    _ = rc(torch.flip(v, (1,)), t, c, j)
    img1.sum().backward()
    t2 = t.clone().requires_grad_(True)
    rc(v, t2, c, j).sum().backward()
    assert rel_l2(t1.grad.cpu().numpy(), t2.grad.cpu().numpy()) <= 1e-5


def test_shape_errors_raise():
    vol, tf, cams, _ = case_inputs((32, 32, 32), (32, 32), 16, seed=26, views=1, jitter=False)
    rc = _rc(vol, (32, 32), 16, jitter=False)
    with pytest.raises(ValueError, match="spatial shape"):
        rc(torch.rand(1, 16, 32, 32, device=DEV), tf.to(DEV), cams[0].to(DEV))
    with pytest.raises(ValueError, match="tf has shape"):
        rc(vol.to(DEV), torch.rand(4, 8, device=DEV), cams[0].to(DEV))


def test_skip_caches_do_not_go_stale_when_the_allocator_recycles_an_address():
    # the per-macro-cell min/max and the cached volume copy are keyed on pointer + version; a fresh tensor
    # (version 0) at a recycled address must not match.  The caches hold the source's storage, so it cannot be recycled.
    out_shape = (64, 48)
    vol, tf, cams, jit = case_inputs((48, 48, 48), out_shape, 128, seed=41, tf_name="tf1", views=1)
    rc_skip = _rc(vol, out_shape, 128, skip_empty=True)
    rc_ref = _rc(vol, out_shape, 128, skip_empty=False)
    t, c, j = tf.to(DEV), cams[0].to(DEV), jit[0].to(DEV)
    a = vol.to(DEV)
    ptr_a = a.data_ptr()
    img_a = rc_skip(a, t, c, j)
    assert torch.equal(img_a, rc_ref(a, t, c, j))
    del a
    b = torch.flip(vol, (1, 2, 3)).contiguous().to(DEV)      # a different volume, same shape/dtype; a fresh tensor has version 0
    img_b = rc_skip(b, t, c, j)
    assert torch.equal(img_b, rc_ref(b, t, c, j))
    assert not torch.equal(img_a, img_b)
    assert torch.equal(rc_skip.vr.last_K, rc_ref.vr.last_K)
    # the cached keys kept the first volume's storage alive, so the allocator could not hand its address to `b`
    assert b.data_ptr() != ptr_a


def test_momentum_step_on_the_volume_invalidates_the_caches():
    # dr_momentum_step writes through a raw pointer; MomentumSGD must bump the version counter
    from differender_b200 import MomentumSGD
    out_shape = (64, 48)
    vol, tf, cams, jit = case_inputs((48, 48, 48), out_shape, 128, seed=42, tf_name="tf1", views=1)
    rc_skip = _rc(vol, out_shape, 128, skip_empty=True)
    rc_ref = _rc(vol, out_shape, 128, skip_empty=False)
    t, c, j = tf.to(DEV), cams[0].to(DEV), jit[0].to(DEV)
    v = vol.to(DEV).contiguous()
    assert torch.equal(rc_skip(v, t, c, j), rc_ref(v, t, c, j))
    ver = v._version
    opt = MomentumSGD(v, lr=1.0, momentum=0.0, max_grad=1.0, lo=0.0, hi=1.0)
    opt.step(torch.where(v > 0.3, -0.3, 0.3).to(v))          # moves many voxels across TF bumps
    assert v._version > ver
    assert torch.equal(rc_skip(v, t, c, j), rc_ref(v, t, c, j))
    assert torch.equal(rc_skip.vr.last_K, rc_ref.vr.last_K)
    with pytest.raises(RuntimeError, match="gradient must be a CUDA tensor"):
        opt.step(torch.zeros(7, device=DEV))
    with pytest.raises(RuntimeError, match="gradient must be a CUDA tensor"):
        opt.step(torch.zeros(v.shape))


def test_inplace_edit_between_forward_and_backward_is_detected():
    # tf / look_from / jitter alias the caller's tensors when already fp32 + contiguous
    from differender_b200 import RaycastFunction
    out_shape = (32, 24)
    vol, tf, cams, jit = case_inputs((32, 32, 32), out_shape, 16, seed=43, views=1)
    rc = _rc(vol, out_shape, 16)
    _, _, vol_in, _, lf_in = rc._determine_batch(vol.to(DEV), tf.to(DEV), cams[0].to(DEV))
    tf_r4 = tf.to(DEV).t().contiguous().requires_grad_(True)                    # [R,4] fp32 contiguous: passed through as is
    raw = RaycastFunction.apply(rc.vr, vol_in, tf_r4, lf_in, 1.0, (False, 0), True, jit[0].to(DEV))
    with torch.no_grad():
        tf_r4.clamp_(0.2, 0.8)
    with pytest.raises(RuntimeError, match="inplace"):       # autograd's saved-tensor version check (either of its messages)
        raw.sum().backward()


def test_cell_major_copy_is_cached_per_volume_version():
    out_shape = (32, 24)
    vol, tf, cams, jit = case_inputs((32, 32, 32), out_shape, 16, seed=44, views=1)
    rc = _rc(vol, out_shape, 16, layout="cell8")
    v = vol.to(DEV)
    lin = v.reshape(1, 32, 32, 32)
    a = rc.vr.brick(lin)
    assert rc.vr.brick(v.reshape(1, 32, 32, 32)) is a        # another view of the same unchanged storage
    v.mul_(0.5)
    b = rc.vr.brick(lin)
    assert b is not a and torch.equal(b[0, :, 0].reshape(32, 32, 32), lin[0])
    rc.vr.forget_volume()
    assert rc.vr.brick(lin) is not b


def test_reference_style_volume_optimisation_loop():
    # examples/test_opt_tf.py:63-88 in miniature: targets from raycast_nondiff of a ground-truth volume, DSSIM + MSE loss, AdamW on
    # the volume, clamp_(0, 1) -- the loss must go down and everything must stay finite
    from differender_b200.losses import mse_dssim_loss
    out_shape = (64, 48)
    vol, tf, cams, jit = case_inputs((32, 32, 32), out_shape, 64, seed=61, tf_name="tf1", views=3)
    rc = _rc(vol, out_shape, 64, sampling_rate=1.0)
    t, c = tf.to(DEV), cams.to(DEV)
    with torch.no_grad():
        gt = rc.raycast_nondiff(vol.to(DEV), t, c, sampling_rate=8.0)
    v = (0.5 * vol + 0.25).to(DEV).requires_grad_(True)
    opt = torch.optim.AdamW([v], lr=2e-2, weight_decay=0)
    losses = []
    for _ in range(6):
        opt.zero_grad()
        loss = mse_dssim_loss(rc(v, t, c), gt)
        loss.backward()
        assert torch.isfinite(v.grad).all()
        opt.step()
        with torch.no_grad():
            v.clamp_(0.0, 1.0)
        losses.append(loss.item())
    assert losses[-1] < losses[0] and all(np.isfinite(losses))
