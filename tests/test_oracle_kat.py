"""Known-answer tests of the oracle that follow directly from the reference's formulas."""
import math

import numpy as np
import torch

from oracle import cpu_oracle as co


def _const_case(val=0.5, alpha=0.05, rgb=(0.9, 0.5, 0.2), N=16, R=8):
    vol = np.full((N, N, N), val, np.float32)
    tf = np.zeros((4, R), np.float32)
    tf[0], tf[1], tf[2], tf[3] = rgb[0], rgb[1], rgb[2], alpha
    return vol, tf


def test_constant_volume_closed_form():
    # flat data: normal is 0/0; the reference launders the NaN through max(.,0) so the Phong factor is the ambient
    # 0.4 (volume_raycaster.py:291-298).  pixel = 0.4*rgb*(1-(1-o)^K), alpha = 1-(1-o)^K, K = first count with alpha >= 0.99
    vol, tf = _const_case()
    cam = np.array([0.0, 0.7, 2.5], np.float32)
    img, K, n = co.forward(vol, tf, cam, (16, 16), max_samples=4096, return_counts=True)
    o = np.float32(0.05)
    hit = n > 0
    assert hit.any()
    for (r, c) in [(8, 8), (7, 9), (3, 12)]:
        if not hit[r, c]:
            continue
        k = K[r, c]
        a = 1.0 - (1.0 - float(o)) ** k
        assert abs(img[3, r, c] - a) < 5e-6
        for ch, col in enumerate((0.9, 0.5, 0.2)):
            assert abs(img[ch, r, c] - 0.4 * col * a) < 5e-6
        # K is the first count reaching 0.99 or the whole ray
        if k < n[r, c]:
            assert 1.0 - (1.0 - float(o)) ** (k - 1) < 0.99 <= a + 1e-6
    assert np.all(img[:, ~hit] == 0.0)


def test_sample_count_formula_centre_ray():
    # n = floor(sr * (tmax - tmin) * ||res - 1||) + 1  (:248-253); a camera on the z axis looks through the centre
    # pixel pair along -z: tmin = dist-1, tmax = dist+1 up to the half-pixel offset.
    N = 32
    vol, tf = _const_case(N=N)
    cam = np.array([0.0, 0.0, 3.0], np.float32)
    for sr in (1.0, 0.5, 2.0):
        _, K, n = co.forward(vol, tf, cam, (64, 64), sampling_rate=sr, max_samples=1 << 20, return_counts=True)
        diag = math.sqrt(3.0) * (N - 1)
        expect = math.floor(sr * 2.0 * diag) + 1
        assert abs(int(n[32, 32]) - expect) <= 1
        assert int(n.max()) <= math.floor(sr * 2.0 * math.sqrt(3.0) * diag) + 1


def test_miss_rays_are_zero_and_have_no_gradient():
    vol, tf = _const_case()
    cam = np.array([0.0, 0.7, 2.5], np.float32)
    img, K, n = co.forward(vol, tf, cam, (24, 24), return_counts=True)
    assert (n == 0).any()
    go = np.zeros_like(img)
    go[:, n == 0] = 1.0                      # gradient only on rays that miss the box
    gv, gt = co.backward(vol, tf, cam, go, (24, 24))
    assert np.all(gv == 0) and np.all(gt == 0)


def test_linear_ramp_normal_is_x_axis():
    # v = a + b*x: trilinear interpolation reproduces it exactly, so the central difference points along +x and
    # the diffuse term is 0.8*max(l.x, 0) with l the direction from the light (cam + (0,1,0)) to the sample (:281-292).
    N = 24
    x = np.linspace(0.2, 0.8, N, dtype=np.float32)
    vol = np.broadcast_to(x[None, None, :], (N, N, N)).copy()
    R = 4
    tf = np.zeros((4, R), np.float32); tf[:3] = 1.0; tf[3] = 1.0      # opaque white: one sample decides the pixel
    cam = np.array([2.5, 0.0, 0.0], np.float32)                       # on the +x axis, looking down -x
    img, K, n = co.forward(vol, tf, cam, (8, 8), return_counts=True)
    hit = n > 0
    assert np.all(K[hit] == 1)
    # l = normalize(pos - light) has l.x < 0 for a camera at +x, so N.l < 0: diffuse 0; r = l - 2(N.l)N flips x,
    # r.(-dir) is then large -> specular.  Pixel must equal k = min(1, 0.3*rdv^32 + 0.4).
    assert np.all(img[0][hit] >= 0.4 - 1e-6) and np.all(img[0][hit] <= 1.0 + 1e-6)
    cam2 = np.array([-2.5, 0.0, 0.0], np.float32)                     # from -x: N.l > 0 -> diffuse lights up
    img2, K2, n2 = co.forward(vol, tf, cam2, (8, 8), return_counts=True)
    c = (4, 4)
    # analytic value for the centre-ish pixel from -x
    from oracle import torch_ref as tr
    ref = tr.render(torch.tensor(vol), torch.tensor(tf), torch.tensor(cam2), (8, 8), dtype=torch.float64).numpy()
    assert np.abs(ref - img2).max() < 1e-4
    assert img2[0][c] > 0.4 + 0.1                                      # diffuse contribution present


def test_alpha_one_sr_one_gradient_is_finite():
    # H8: TF alpha == 1 at sr == 1 has d/da (1-a)^(1/sr) = 1: finite gradient
    vol, tf = _const_case(alpha=1.0)
    vol += np.random.default_rng(0).uniform(0, 1e-2, vol.shape).astype(np.float32)
    cam = np.array([0.0, 0.7, 2.5], np.float32)
    img = co.forward(vol, tf, cam, (8, 8))
    gv, gt = co.backward(vol, tf, cam, np.ones_like(img), (8, 8))
    assert np.isfinite(gv).all() and np.isfinite(gt).all()
