"""Caller-side steps next to the march: oracle known answers on CPU, CUDA parity on the GPU."""
import numpy as np
import pytest
import torch

from helpers import GRAD_TOL, case_inputs, oracle_backward_views, oracle_forward_views, rel_l2
from oracle import aux_ref


def test_momentum_oracle_known_answers():
    p = np.array([0.5, 0.01, 0.2, 0.9], np.float32); g = np.array([1.0, 0.05, -3.0, 0.0], np.float32); m = np.zeros(4, np.float32)
    p1, m1 = aux_ref.momentum_step(p, g, m, lr=0.1, gamma=0.9, max_grad=0.1)
    assert np.allclose(m1, [0.01, 0.005, -0.01, 0.0]) and np.allclose(p1, [0.49, 0.005, 0.21, 0.9])
    p2, m2 = aux_ref.momentum_step(p1, g, m1, lr=0.1, gamma=0.9, max_grad=0.1)
    assert np.allclose(m2, [0.019, 0.0095, -0.019, 0.0]) and np.allclose(p2, [0.471, 0.0, 0.229, 0.9], atol=1e-7)   # max(tf, 0)
    p3, _ = aux_ref.momentum_step(np.array([0.99], np.float32), np.array([-1.0], np.float32), np.zeros(1, np.float32), 1.0, 0.0, 0.1, 0.0, 1.0)
    assert p3[0] == 1.0                                                                                           # clamp_(0, 1)


def test_ingest_oracle_known_answers():
    raw = np.arange(2 * 3 * 4, dtype=np.uint8).reshape(2, 3, 4)
    v = aux_ref.ingest_u8(raw)
    assert v.shape == (3, 2, 4) and v[1, 0, 2] == np.float32(raw[0, 1, 2]) / np.float32(255) and v.dtype == np.float32
    assert aux_ref.ingest_u8(np.array([[[255, 0]]], np.uint8), False).tolist() == [[[1.0, 0.0]]]


@pytest.mark.gpu
def test_momentum_step_matches_oracle_bit_for_bit():
    from differender_b200 import MomentumSGD
    g = torch.Generator().manual_seed(0)
    p = torch.rand(4, 128, generator=g); grads = [torch.randn(4, 128, generator=g) * 0.2 for _ in range(5)]
    pc = p.cuda().contiguous()
    opt = MomentumSGD(pc, lr=0.1, momentum=0.9, max_grad=0.1, lr_decay=0.99)
    pn, mn, lr = p.numpy().copy(), np.zeros((4, 128), np.float32), 0.1
    for gr in grads:
        opt.step(gr.cuda())
        pn, mn = aux_ref.momentum_step(pn, gr.numpy(), mn, lr, 0.9, 0.1)
        lr *= 0.99
    assert np.array_equal(pc.cpu().numpy(), pn) and np.array_equal(opt.state.cpu().numpy(), mn)
    vol = torch.rand(1, 8, 8, 8).cuda()
    MomentumSGD(vol, lr=1.0, momentum=0.0, max_grad=10.0, hi=1.0).step(torch.full((1, 8, 8, 8), -5.0).cuda())
    assert vol.max().item() == 1.0 and vol.min().item() == 1.0


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
def test_ingest_u8_matches_oracle(dtype):
    from differender_b200.utils import volume_from_raw_u8
    raw = np.random.default_rng(1).integers(0, 256, (20, 12, 28), dtype=np.uint8)        # [Z][Y][X] on disk
    ref = aux_ref.ingest_u8(raw, True)
    vol = volume_from_raw_u8(raw, swap_axes01=True, dtype=dtype)
    assert vol.shape == (1, 12, 20, 28) and vol.dtype == dtype
    want = torch.tensor(ref).to(dtype)                                                    # fp16: the fp32 quotient rounded once
    assert torch.equal(vol[0].cpu(), want)
    assert torch.equal(volume_from_raw_u8(raw.reshape(-1), shape=(20, 12, 28), swap_axes01=False)[0].cpu(), torch.tensor(aux_ref.ingest_u8(raw, False)))


@pytest.mark.gpu
def test_fused_mse_matches_unfused_and_oracle():
    from differender_b200 import Raycaster
    out_shape = (48, 40)
    vol, tf, cams, jit = case_inputs((36, 36, 36), out_shape, 64, seed=31, tf_name="tf1", views=2)
    target = torch.rand(2, 4, 40, 48, generator=torch.Generator().manual_seed(3))
    rc = Raycaster((36, 36, 36), out_shape, 64, max_samples=2048)
    dev = "cuda:0"
    v1 = vol.to(dev).requires_grad_(True); t1 = tf.to(dev).requires_grad_(True)
    loss1, img1 = rc.mse_loss(v1, t1, cams.to(dev), target.to(dev), jit.to(dev))
    loss1.backward()
    v2 = vol.to(dev).requires_grad_(True); t2 = tf.to(dev).requires_grad_(True)
    img2 = rc(v2, t2, cams.to(dev), jit.to(dev))
    loss2 = torch.nn.functional.mse_loss(img2, target.to(dev))
    loss2.backward()
    assert torch.equal(img1, img2.detach()) and not img1.requires_grad
    assert abs(loss1.item() - loss2.item()) <= 1e-6 * abs(loss2.item())
    assert rel_l2(v1.grad.cpu().numpy(), v2.grad.cpu().numpy()) <= 1e-5 and rel_l2(t1.grad.cpu().numpy(), t2.grad.cpu().numpy()) <= 1e-5
    ref, _, _ = oracle_forward_views(vol, tf, cams, out_shape, jit, max_samples=2048)
    lref, go = aux_ref.mse(ref, target.numpy())
    gv, gt = oracle_backward_views(vol, tf, cams, go.astype(np.float32), out_shape, jit, max_samples=2048)
    assert abs(loss1.item() - lref) <= 1e-5 * lref
    assert rel_l2(v1.grad[0].cpu().numpy(), gv) <= GRAD_TOL and rel_l2(t1.grad.cpu().numpy(), gt) <= GRAD_TOL
    # a scaled upstream gradient scales the result
    v3 = vol.to(dev).requires_grad_(True)
    (3.0 * rc.mse_loss(v3, tf.to(dev), cams.to(dev), target.to(dev), jit.to(dev))[0]).backward()
    assert rel_l2(v3.grad.cpu().numpy(), 3.0 * v2.grad.cpu().numpy()) <= 1e-5


def test_ssim_and_the_reference_training_loss():
    # reference examples/test_opt_tf.py:70-72: nan_to_num(1 - ssim(res, gt, data_range=1, nonnegative_ssim=True)) + mse_loss(res, gt)
    from differender_b200.losses import dssim_loss, mse_dssim_loss, ssim
    g = torch.Generator().manual_seed(0)
    a = torch.rand(2, 4, 40, 48, generator=g, dtype=torch.float64)
    b = (a + 0.1 * torch.randn(2, 4, 40, 48, generator=g, dtype=torch.float64)).clamp(0, 1)
    assert abs(ssim(a, a).item() - 1.0) < 1e-12
    ref = aux_ref.ssim(a.numpy(), b.numpy(), nonnegative=True)
    assert abs(ssim(a, b, nonnegative_ssim=True).item() - ref) < 1e-10 and 0.0 < ref < 1.0
    assert abs(dssim_loss(a, b).item() - (1.0 - ref)) < 1e-10
    mref, _ = aux_ref.mse(a.numpy(), b.numpy())
    assert abs(mse_dssim_loss(a, b).item() - (1.0 - ref + mref)) < 1e-10
    # differentiable (the gradient reaches the march through Raycaster.forward's ordinary backward)
    x = a.clone().requires_grad_(True)
    mse_dssim_loss(x, b).backward()
    assert torch.isfinite(x.grad).all() and x.grad.abs().max() > 0
    with pytest.raises(ValueError):
        ssim(a[0], b[0])
