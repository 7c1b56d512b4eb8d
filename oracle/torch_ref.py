"""Vectorised PyTorch restatement of the reference ray-march, differentiated by torch.autograd.

TEST INFRASTRUCTURE, NOT PRODUCT CODE (only tests/ and fixture generators import it).

Second, independent oracle: same semantics as oracle/cpu_ref.c (both restate
/root/reference/differender/volume_raycaster.py), but the gradients here come from torch.autograd
instead of a hand-written adjoint, and the arithmetic can run in float64.  It pins the hand-derived
backward of cpu_ref.c (and through it the CUDA backward) and generates tests/golden/*.npz.

PARITY UNPINNED against the real reference: Taichi / taichi_glsl are not installable here (no
network); see the header of cpu_ref.c.

Line references: volume_raycaster.py in the reference.
"""
import math

import numpy as np
import torch


def _normalized(v):
    # Taichi Vector.normalized(eps=0): v * (1 / norm)
    return v * (1.0 / torch.sqrt((v * v).sum(-1, keepdim=True)))


def _mix(a, b, t):
    # taichi_glsl.mix
    return a * (1.0 - t) + b * t


def _fmax0(x):
    # ti.max(x, 0.0): returns the non-NaN operand; gradient passes iff x > 0
    return torch.where(x > 0, x, torch.zeros_like(x))


class _Consts:
    def __init__(self, vol_shape_dhw, output_shape, R, sr, fov, near, dtype):
        D, Hv, Wv = vol_shape_dhw
        w, h = output_shape
        f32 = np.float32
        self.dim = (Wv, D, Hv)                                  # Taichi (X,Y,Z)  :481
        near_h = 2.0 * math.tan(math.radians(fov)) * near       # :146 (tan of the full angle)
        near_w = near_h * (w / h)                               # :147, :75
        # Python scalars captured by the kernels are f32 constants on the device
        self.near = float(f32(near)); self.near_h = float(f32(near_h)); self.near_w = float(f32(near_w))
        self.scale = [float(f32(d - 1.0 - 1e-4)) for d in self.dim]           # :165
        self.vol_diag = float(f32(math.sqrt(sum((d - 1.0) ** 2 for d in self.dim))))   # :248-249
        self.tf_len = float(R - 1)                              # :215
        self.sr = float(f32(sr)); self.inv_sr = float(f32(1.0 / sr))
        self.w, self.h, self.R = w, h, R
        self.dtype = dtype


def _rays(c, cam, jitter_raw, has_jitter):
    """compute_entry_exit :221-259 for all pixels; raw order [w,h]."""
    dt = c.dtype
    ii = torch.arange(c.w, dtype=dt).view(-1, 1).expand(c.w, c.h)
    jj = torch.arange(c.h, dtype=dt).view(1, -1).expand(c.w, c.h)
    x = (ii + 0.5) / float(c.w)
    y = (jj + 0.5) / float(c.h)
    cam = cam.to(dt)
    view = -cam * (1.0 / torch.sqrt((cam * cam).sum()))        # (-look_from).normalized() :233
    up0 = torch.tensor([0.0, 1.0, 0.0], dtype=dt)
    right = _normalized(torch.linalg.cross(view, up0))
    up = _normalized(torch.linalg.cross(right, view))
    u = (x - 0.5).unsqueeze(-1); v = (y - 0.5).unsqueeze(-1)
    near_m = cam + c.near * view
    near_pos = near_m + (u * c.near_w) * right + (v * c.near_h) * up
    vd = _normalized(near_pos - cam)                             # :148-151
    inv = 1.0 / vd                                               # :41
    ta = (-1.0 - cam) * inv
    tb = (1.0 - cam) * inv
    tlo = torch.minimum(ta, tb); thi = torch.maximum(ta, tb)
    tmin = torch.maximum(torch.maximum(tlo[..., 0], tlo[..., 1]), tlo[..., 2])
    tmax = torch.minimum(torch.minimum(thi[..., 0], thi[..., 1]), thi[..., 2])
    hit = ~((tmax < 0) | (tmin > tmax))                          # :52
    ray_len = tmax - tmin
    nf = torch.where(hit, torch.floor(c.sr * ray_len * c.vol_diag) + 1.0, torch.zeros_like(ray_len))   # :251-253
    n = nf.to(torch.int64)
    entry = tmin
    if has_jitter:
        entry = torch.where(n > 0, tmin + jitter_raw.to(dt) * ray_len / nf.clamp(min=1.0), tmin)       # :254-255
    return cam, vd, entry, tmax, n


def _trilinear(c, volflat, pos):
    """sample_volume_trilinear :153-189; volflat is the torch (D,H,W) volume flattened: idx = (y*Z + z)*X + x."""
    X, Y, Z = c.dim
    lo = []; hi = []; fr = []
    for a, dim in enumerate(c.dim):
        p = torch.clamp(0.5 * pos[..., a] + 0.5, 0.0, 1.0) * c.scale[a]
        p = torch.clamp(p, min=0.0)
        l = torch.floor(p)
        fr.append(p - l)
        li = l.to(torch.int64)
        lo.append(li); hi.append(torch.clamp(li + 1, max=dim - 1))

    def at(xi, yi, zi):
        return volflat[(yi * Z + zi) * X + xi]

    fx, fy, fz = fr
    a = _mix(at(lo[0], lo[1], lo[2]), at(hi[0], lo[1], lo[2]), fx)
    b = _mix(at(lo[0], hi[1], lo[2]), at(hi[0], hi[1], lo[2]), fx)
    zl = _mix(a, b, fy)
    a = _mix(at(lo[0], lo[1], hi[2]), at(hi[0], lo[1], hi[2]), fx)
    b = _mix(at(lo[0], hi[1], hi[2]), at(hi[0], hi[1], hi[2]), fx)
    zh = _mix(a, b, fy)
    return _mix(zl, zh, fz)


def _shade(c, volflat, tf_r4, cam, vd, pos, nondiff):
    """Body of raycast :281-299 for a set of rays.  Returns (C [.,4], alpha of the TF sample)."""
    dt = c.dtype
    I = _trilinear(c, volflat, pos)
    x = I * c.tf_len
    x = torch.where(x > 0, x, torch.zeros_like(x))                # low_high_frac: ti.max(x, 0)
    l = torch.floor(x)
    f = (x - l).unsqueeze(-1)
    lo = torch.clamp(l.to(torch.int64), max=c.R - 1)
    hi = torch.clamp(lo + 1, max=c.R - 1)
    col = _mix(tf_r4[lo], tf_r4[hi], f)                           # :217-219
    alpha = col[..., 3]
    base = 1.0 - alpha
    if c.inv_sr == 1.0:
        o = 1.0 - base
    else:
        o = 1.0 - torch.pow(torch.clamp(base, min=1e-12), c.inv_sr)   # :284-285 (H8 clamp)
    delta = 1e-3                                                   # :193
    g = []
    for a in range(3):
        e = torch.zeros(3, dtype=dt); e[a] = delta
        g.append(_trilinear(c, volflat, pos + e) - _trilinear(c, volflat, pos - e))
    g = torch.stack(g, -1)
    glen2 = (g * g).sum(-1)
    flat = glen2 == 0                                              # H4
    gs = torch.where(flat.unsqueeze(-1), torch.ones_like(g), g)
    N = _normalized(gs)
    lp = cam + torch.tensor([0.0, 1.0, 0.0], dtype=dt)            # :281
    ld = _normalized(pos - lp)                                     # :288-290
    nl = (N * ld).sum(-1)
    ndl = torch.where(flat, torch.zeros_like(nl), _fmax0(nl))      # :291
    r = ld - (2.0 * nl).unsqueeze(-1) * N                          # tl.reflect
    rv = (r * (-vd)).sum(-1)
    rdv = torch.where(flat, torch.zeros_like(rv), _fmax0(rv))      # :295
    kraw = 0.8 * ndl + 0.3 * rdv ** 32 + 0.4                       # :292-298
    k = kraw if nondiff else torch.where(kraw > 1.0, torch.ones_like(kraw), kraw)
    rgb = (k.unsqueeze(-1) * col[..., :3]) * o.unsqueeze(-1)       # :297-299
    C = torch.cat([rgb, o.unsqueeze(-1)], -1)
    return C, alpha


def render(volume, tf, look_from, output_shape, sampling_rate=1.0, max_samples=512, fov=30.0, near=0.1,
           jitter=None, nondiff=False, dtype=torch.float64, return_counts=False):
    """One view.  volume (D,H,W)/(1,D,H,W), tf (4,R), look_from (3,), jitter (H,W) in image orientation or None.
    Returns the image (4,H,W) in `dtype` (differentiable w.r.t. volume and tf) [and K, n as (H,W) int64]."""
    vol = volume.reshape(volume.shape[-3:]).to(dtype)
    c = _Consts(tuple(vol.shape), output_shape, tf.shape[-1], sampling_rate, fov, near, dtype)
    volflat = vol.reshape(-1)
    tf_r4 = tf.to(dtype).permute(1, 0)                             # :571
    jr = None
    if jitter is not None:
        jr = torch.flip(jitter.to(dtype).permute(1, 0), (1,))      # image (H,W) -> raw (w,h): J[H-1-j, i]
    cam, vd, entry, exit_, n = _rays(c, torch.as_tensor(look_from), jr, jitter is not None)
    w, h = c.w, c.h
    A = torch.zeros(w, h, 4, dtype=dtype)
    K = torch.zeros(w, h, dtype=torch.int64)
    nmax = int(n.max().item()) if n.numel() else 0
    nf = n.to(dtype)
    ray_len = exit_ - entry                                        # :272 (from the jittered entry)
    t0 = entry + 0.5 * ray_len / nf.clamp(min=1.0)                 # :273-275
    for s in range(nmax):
        active = (A[..., 3].detach() < 0.99) & (s < n)
        if not nondiff:
            active = active & (s < max_samples)                    # :267-269
        if not bool(active.any()):
            break
        idx = active.nonzero(as_tuple=True)
        ratio = torch.where(n[idx] > 1, float(s) * (1.0 / (nf[idx] - 1.0).clamp(min=1.0)), torch.zeros_like(nf[idx]))  # H3; s * fl(1/(n-1))
        t = _mix(t0[idx], exit_[idx], ratio)                       # :277-280
        pos = cam + t.unsqueeze(-1) * vd[idx]
        C, alpha = _shade(c, volflat, tf_r4, cam, vd[idx], pos, nondiff)
        Aa = A[idx]
        newA = (1.0 - Aa[..., 3:4]) * C + Aa                       # :300-302
        if nondiff:
            take = alpha.detach() > 1e-3                           # :334
            newA = torch.where(take.unsqueeze(-1), newA, Aa)
            K = K.index_put(idx, K[idx] + take.to(torch.int64))
        else:
            K = K.index_put(idx, K[idx] + 1)
        A = A.index_put(idx, newA)
    if nondiff:
        A = torch.clamp(A, max=1.0)                                # :358
    img = torch.flip(A, (1,)).permute(2, 1, 0).contiguous()        # :543-548
    if return_counts:
        to_img = lambda q: torch.flip(q, (1,)).permute(1, 0).contiguous()
        return img, to_img(K), to_img(n)
    return img
