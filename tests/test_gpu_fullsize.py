"""BASELINE.json's full sizes on the GPU.  The oracle cannot finish these in seconds, so the checks are size-independent
properties of the path: determinism, bounds, batching invariance, linearity of the backward in grad_out, plus one oracle
comparison on a reduced ray count through the full-size volume."""
import numpy as np
import pytest
import torch

from helpers import GRAD_TOL, RGBA_TOL, rel_l2

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _setup(n, res, views, dtype=torch.float32, M=2048):
    from differender_b200 import VolumeRaycaster
    from differender_b200.synthetic import make_cameras, make_jitter, make_tf, make_volume
    vol = make_volume(n, device=DEV, dtype=dtype)
    tf = make_tf("tf1", 128, device=DEV)
    cams = make_cameras(16, device=DEV)[:views].contiguous()
    jit = make_jitter(views, res[1], res[0], device=DEV)
    vr = VolumeRaycaster((n, n, n), res, max_samples=M, tf_resolution=128)
    bricked = vr.brick(vol.reshape(1, n, n, n))
    return vr, vol, tf, tf.t().contiguous()[None], cams, jit, bricked


def _props(vr, bricked, tf_r4, cams, jit, M):
    out, K, Tp = vr.march(bricked, tf_r4, cams, 1.0, jit)
    out2, K2, _ = vr.march(bricked, tf_r4, cams, 1.0, jit)
    assert torch.equal(out, out2) and torch.equal(K, K2)                         # forward is deterministic (no atomics)
    assert torch.isfinite(out).all() and out.min().item() >= 0.0
    assert 0 <= K.min().item() and K.max().item() <= M
    assert out[:, 3].max().item() <= 1.0 + 1e-5
    term = out[:, 3] >= 0.99
    assert term.any() and (Tp[term] > 0.01 - 1e-6).all()                         # a terminated ray was still active before its last sample
    assert (out.permute(1, 0, 2, 3)[:, K == 0] == 0).all()                                       # rays without an active sample are empty
    one, K1, _ = vr.march(bricked, tf_r4, cams[:1].contiguous(), 1.0, jit[:1].contiguous())
    assert torch.equal(one[0], out[0]) and torch.equal(K1[0], K[0])              # batching does not change a view
    return out, K, Tp


def _backward_linearity(vr, bricked, tf_r4, cams, jit, out, K, Tp):
    g = torch.Generator(device=DEV).manual_seed(3)
    g1 = torch.randn(out.shape, generator=g, device=DEV); g2 = torch.randn(out.shape, generator=g, device=DEV)
    a = (bricked, tf_r4, cams, 1.0, jit)
    v1, t1 = vr.march_backward(*a, g1, out, K, Tp, True, True)
    v2, t2 = vr.march_backward(*a, g2, out, K, Tp, True, True)
    v3, t3 = vr.march_backward(*a, (0.5 * g1 - 2.0 * g2).contiguous(), out, K, Tp, True, True)
    assert torch.isfinite(v3).all() and torch.isfinite(t3).all()
    assert rel_l2((0.5 * v1 - 2.0 * v2).cpu().numpy(), v3.cpu().numpy()) <= 1e-4
    assert rel_l2((0.5 * t1 - 2.0 * t2).cpu().numpy(), t3.cpu().numpy()) <= 1e-4
    z, zt = vr.march_backward(*a, torch.zeros_like(g1), out, K, Tp, True, True)
    assert (z == 0).all() and (zt == 0).all()
    # per-view gradients add up to the batch gradient (one shared cell-major buffer, no per-view clones)
    s_v = torch.zeros_like(v1); s_t = torch.zeros_like(t1)
    for i in range(cams.shape[0]):
        sl = slice(i, i + 1)
        vi, ti = vr.march_backward(bricked, tf_r4, cams[sl].contiguous(), 1.0, jit[sl].contiguous(), g1[sl].contiguous(),
                                   out[sl].contiguous(), K[sl].contiguous(), Tp[sl].contiguous(), True, True)
        s_v += vi; s_t += ti
    assert rel_l2(s_v.cpu().numpy(), v1.cpu().numpy()) <= 1e-4 and rel_l2(s_t.cpu().numpy(), t1.cpu().numpy()) <= 1e-4


def test_c3_256cube_1024sq():
    vr, vol, tf, tf_r4, cams, jit, bricked = _setup(256, (1024, 1024), 2, M=2048)
    out, K, Tp = _props(vr, bricked, tf_r4, cams, jit, 2048)
    _backward_linearity(vr, bricked, tf_r4, cams, jit, out, K, Tp)


def test_c3_volume_against_oracle_on_reduced_ray_count():
    # the full 256^3 volume, 128x128 rays: the oracle finishes in seconds
    from oracle import cpu_oracle as co
    vr, vol, tf, tf_r4, cams, jit, bricked = _setup(256, (128, 128), 1, M=2048)
    out, K, Tp = vr.march(bricked, tf_r4, cams, 1.0, jit)
    v, t, c, j = vol.cpu().numpy(), tf.cpu().numpy(), cams[0].cpu().numpy(), jit[0].cpu().numpy()
    ref, Kr, _ = co.forward(v, t, c, (128, 128), max_samples=2048, jitter=j, return_counts=True)
    same = K[0].cpu().numpy() == Kr
    assert (~same).mean() <= 1e-4
    assert np.abs(out[0].cpu().numpy() - ref)[:, same].max() <= RGBA_TOL
    go = torch.randn(out.shape, generator=torch.Generator().manual_seed(8))
    gv, gt = vr.march_backward(bricked, tf_r4, cams, 1.0, jit, go.to(DEV), out, K, Tp, True, True)
    gvr, gtr = co.backward(v, t, c, go[0].numpy(), (128, 128), max_samples=2048, jitter=j)
    assert rel_l2(gv[0].cpu().numpy(), gvr) <= GRAD_TOL and rel_l2(gt[0].cpu().numpy().T, gtr) <= GRAD_TOL


def test_c4_512cube_1024sq():
    vr, vol, tf, tf_r4, cams, jit, bricked = _setup(512, (1024, 1024), 2, M=4096)
    out, K, Tp = _props(vr, bricked, tf_r4, cams, jit, 4096)
    _backward_linearity(vr, bricked, tf_r4, cams, jit, out, K, Tp)


def test_c5_1024cube_fp16_2048sq():
    vr, vol, tf, tf_r4, cams, jit, bricked = _setup(1024, (2048, 2048), 1, dtype=torch.float16, M=8192)
    assert bricked.dtype == torch.float16 and bricked.numel() == 8 * 1024 ** 3       # `auto` = cell-major copy (16 GiB)
    out, K, Tp = vr.march(bricked, tf_r4, cams, 1.0, jit)
    assert torch.isfinite(out).all() and K.max().item() <= 8192 and out[:, 3].max().item() <= 1.0 + 1e-5
    g1 = torch.randn(out.shape, generator=torch.Generator(device=DEV).manual_seed(1), device=DEV)
    # 32 GiB cell-major gradient buffer, gathered to the 4 GiB linear gradient
    gv, gt = vr.march_backward(bricked, tf_r4, cams, 1.0, jit, g1, out, K, Tp, True, True)
    assert gv.shape == (1, 1024, 1024, 1024) and torch.isfinite(gt).all()
    gv2, gt2 = vr.march_backward(bricked, tf_r4, cams, 1.0, jit, (2.0 * g1).contiguous(), out, K, Tp, False, True)
    assert gv2 is None and rel_l2(gt2.cpu().numpy(), 2.0 * gt.cpu().numpy()) <= 1e-4
    assert torch.isfinite(gv[0, ::64]).all() and gv.abs().max().item() > 0


def test_c5_skip_grid_is_bit_identical_at_full_size():
    # the fp16 SKIP forward at 1024^3 / 2048^2 (kSkipMargin was argued from an fp32 position-error estimate): image, K and
    # Tprev must equal the forward that marches every sample, bit for bit
    vr, vol, tf, tf_r4, cams, jit, bricked = _setup(1024, (2048, 2048), 1, dtype=torch.float16, M=8192)
    a = vr.march(bricked, tf_r4, cams, 1.0, jit, skip=True)
    grid = vr.skip_grid(vr.desc(1, 1, 1, 1, 0, 1.0), bricked, tf_r4)
    assert grid is not None and int(grid[:4].view(torch.int32)[0]) > 0        # the grid exists and has empty macro-cells
    b = vr.march(bricked, tf_r4, cams, 1.0, jit, skip=False)
    for x, y in zip(a, b):
        assert torch.equal(x, y)
    assert a[1].sum().item() > 0


def test_c5_volume_against_oracle_on_reduced_ray_count():
    # the full 1024^3 fp16 volume through the product-default kernels (fp16 x cell8 x two-neighbour taps x skip grid) on
    # 64x64 rays, against the oracle marching the same fp16-rounded values: image, K, TF gradient and volume gradient
    from oracle import cpu_oracle as co
    n, res = 1024, (64, 64)
    vr, vol, tf, tf_r4, cams, jit, bricked = _setup(n, res, 1, dtype=torch.float16, M=8192)
    assert bricked.dtype == torch.float16 and bricked.ndim == 3
    out, K, Tp = vr.march(bricked, tf_r4, cams, 1.0, jit)
    v = vol.reshape(n, n, n).cpu().numpy().astype(np.float32)                 # 4 GiB on the host
    t, c, j = tf.cpu().numpy(), cams[0].cpu().numpy(), jit[0].cpu().numpy()
    ref, Kr, _ = co.forward(v, t, c, res, max_samples=8192, jitter=j, return_counts=True)
    same = K[0].cpu().numpy() == Kr
    assert (~same).mean() <= 1e-3, f"{(~same).sum()} of {same.size} rays differ in active sample count"
    assert np.abs(out[0].cpu().numpy() - ref)[:, same].max() <= RGBA_TOL
    go = torch.randn(out.shape, generator=torch.Generator().manual_seed(8))
    gv, gt = vr.march_backward(bricked, tf_r4, cams, 1.0, jit, go.to(DEV), out, K, Tp, True, True)
    gvr, gtr = co.backward(v, t, c, go[0].numpy(), res, max_samples=8192, jitter=j)     # gvr: lazily-committed zeros except near the rays
    assert rel_l2(gt[0].cpu().numpy().T, gtr) <= GRAD_TOL
    # the volume gradient is sparse (4096 rays through 1e9 voxels): compare on the oracle's support, and require the GPU
    # gradient to carry (almost) nothing outside it
    idx = np.flatnonzero(gvr.reshape(-1))
    assert idx.size > 10000
    ref_nz = gvr.reshape(-1)[idx]
    got_nz = gv.reshape(-1)[torch.from_numpy(idx).to(DEV)].double().cpu().numpy()
    assert np.linalg.norm(got_nz - ref_nz) / np.linalg.norm(ref_nz) <= GRAD_TOL
    total = float(gv.double().norm().item())
    assert abs(total ** 2 - float(np.linalg.norm(got_nz)) ** 2) <= (GRAD_TOL * total) ** 2
