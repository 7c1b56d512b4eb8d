"""dr_gather_grad (tile kernel) and dr_gather_step (gather + momentum/projection step + cell-major volume refresh fused,
SURVEY 8(f) row 2) against the host restatements: bit for bit."""
import ctypes

import numpy as np
import pytest
import torch

import hostsim_lib as hs
from helpers import case_inputs, rel_l2
from oracle import aux_ref

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _vr(shape, layout="cell8", out_shape=(32, 24), R=16):
    from differender_b200 import VolumeRaycaster
    D, H, W = shape
    return VolumeRaycaster((W, D, H), out_shape, max_samples=2048, tf_resolution=R, layout=layout)


def _random_cells(shape, seed, sparse=True):
    g = np.random.default_rng(seed)
    n = int(np.prod(shape))
    cells = g.standard_normal((n, 8)).astype(np.float32)
    if sparse:
        cells[g.random(n) < 0.5] = 0.0                     # untouched cells, as after a real backward
    return cells


@pytest.mark.parametrize("shape", [(9, 11, 37), (16, 16, 64), (5, 70, 33), (33, 8, 32)])
def test_gather_kernel_matches_gather_voxel_bit_for_bit(shape):
    vr = _vr(shape)
    cells = _random_cells(shape, 1)                         # includes non-zero cells on the last planes (the clamped slots)
    cells[7, 3] = np.nan; cells[11, 0] = np.inf; cells[12, 1] = -np.inf
    ref = hs.gather(cells, shape)
    ref = np.nan_to_num(ref, nan=0.0, posinf=np.finfo(np.float32).max, neginf=np.finfo(np.float32).min)
    got = vr.gather(torch.tensor(cells).reshape(1, -1).to(DEV))
    assert got.shape == (1,) + shape
    assert np.array_equal(got[0].cpu().numpy(), ref)
    # accumulate
    out = torch.ones((1,) + shape, device=DEV)
    from differender_b200 import _lib
    d = vr.desc(1, 1, 1, _lib.VOX_F32, 0, 1.0)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(_lib.load().dr_gather_grad(ctypes.byref(d), _lib.ptr(torch.tensor(cells).to(DEV)), _lib.ptr(out), 1, st), "gather")
    assert np.array_equal(out[0].cpu().numpy(), (np.float32(1.0) + ref).astype(np.float32))


@pytest.mark.parametrize("shape,dtype", [((9, 11, 37), torch.float32), ((16, 24, 64), torch.float32), ((12, 9, 40), torch.float16)])
@pytest.mark.parametrize("from_cells", [True, False])
def test_gather_step_equals_gather_then_momentum_then_expand(shape, dtype, from_cells):
    from differender_b200 import _lib
    vr = _vr(shape)
    g = np.random.default_rng(5)
    p0 = g.random(shape).astype(np.float32); m0 = (0.05 * g.standard_normal(shape)).astype(np.float32)
    cells = _random_cells(shape, 2) * np.float32(0.2)
    cells[:, :][np.arange(0, cells.shape[0], 97)] = np.nan          # nan_to_num inside the fused kernel
    glin = np.nan_to_num(hs.gather(cells, shape), nan=0.0, posinf=np.finfo(np.float32).max, neginf=np.finfo(np.float32).min)
    lr, gamma, mg, lo, hi = 0.3, 0.9, 0.1, 0.0, 1.0
    p_ref, m_ref = aux_ref.momentum_step(p0, glin, m0, lr, gamma, mg, lo, hi)
    rec_ref = hs.expand(p_ref)
    if dtype == torch.float16:
        rec_ref = torch.tensor(rec_ref).half().numpy()
    P = torch.tensor(p0).to(DEV); M = torch.tensor(m0).to(DEV)
    V = torch.full((p0.size, 8), -7.0, dtype=dtype, device=DEV)       # stale copy: every slot must be rewritten
    G = torch.empty(shape, dtype=torch.float32, device=DEV)
    d = vr.desc(1, 1, 1, _lib.VOX_F16 if dtype == torch.float16 else _lib.VOX_F32, 0, 1.0)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    C = torch.tensor(cells).to(DEV) if from_cells else None
    L = None if from_cells else torch.tensor(glin).to(DEV)
    _lib.check(_lib.load().dr_gather_step(ctypes.byref(d), _lib.ptr(C), _lib.ptr(L), _lib.ptr(P), _lib.ptr(M), _lib.ptr(V), _lib.ptr(G),
                                          lr, gamma, mg, lo, hi, st), "dr_gather_step")
    assert np.array_equal(P.cpu().numpy(), p_ref) and np.array_equal(M.cpu().numpy(), m_ref)
    assert np.array_equal(G.cpu().numpy(), glin)
    assert np.array_equal(V.cpu().numpy(), rec_ref)
    with pytest.raises(RuntimeError, match="exactly one"):
        _lib.check(_lib.load().dr_gather_step(ctypes.byref(d), None, None, _lib.ptr(P), _lib.ptr(M), None, None, lr, gamma, mg, lo, hi, st), "x")


def test_fused_volume_sgd_loop_matches_unfused_loop():
    # three optimisation steps: autograd backward + MomentumSGD (gather, step and re-expansion as separate passes) vs
    # FusedVolumeSGD (the gradient stays cell-major, one kernel per step, the cell-major volume copy refreshed in place)
    from differender_b200 import FusedVolumeSGD, MomentumSGD, Raycaster
    out_shape = (48, 40)
    vol, tf, cams, jit = case_inputs((40, 36, 44), out_shape, 64, seed=51, tf_name="tf1", views=2)
    target = torch.rand(2, 4, 40, 48, generator=torch.Generator().manual_seed(3)).to(DEV)
    t, c, j = tf.to(DEV), cams.to(DEV), jit.to(DEV)
    res = []
    for fused in (False, True):
        rc = Raycaster((40, 36, 44), out_shape, 64, max_samples=2048, layout="cell8")
        v = vol.to(DEV).clone().requires_grad_(True)
        opt = FusedVolumeSGD(rc, v, lr=5.0, momentum=0.9, max_grad=0.05, hi=1.0) if fused else \
            MomentumSGD(v, lr=5.0, momentum=0.9, max_grad=0.05, lo=0.0, hi=1.0)
        losses = []
        for _ in range(3):
            loss = torch.nn.functional.mse_loss(rc(v, t, c, j), target)
            loss.backward()
            losses.append(loss.item())
            if fused:
                assert v.grad is None                      # the gradient never left its cell-major form
                opt.step()
                assert rc.vr._cached_copy(v.view(1, 40, 36, 44)) is not None      # the next forward re-uses the refreshed copy
            else:
                opt.step()
                v.grad = None
        res.append((v.detach().clone(), opt.state.clone(), losses, rc))
    (va, ma, la, _), (vb, mb, lb, rcb) = res
    assert la[2] < la[0] and lb[2] < lb[0]
    # float atomics order differs from run to run, so the two loops agree to rounding, not bit for bit
    assert rel_l2(vb.cpu().numpy(), va.cpu().numpy()) <= 1e-5 and rel_l2(mb.cpu().numpy(), ma.cpu().numpy()) <= 1e-3
    assert abs(la[2] - lb[2]) <= 1e-5 * abs(la[2])
    # and the refreshed copy equals a fresh expansion of the final volume
    rc2 = Raycaster((40, 36, 44), out_shape, 64, max_samples=2048, layout="cell8")
    fresh = rc2.vr.brick(vb.view(1, 40, 36, 44))
    assert torch.equal(rcb.vr._copy_cache[2], fresh)
