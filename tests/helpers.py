"""Shared test inputs and comparison helpers."""
import math

import numpy as np
import torch

from differender_b200.synthetic import make_cameras, make_jitter, make_tf, make_volume

# tolerances stated by BASELINE.json north_star
RGBA_TOL = 1e-4        # max-abs
GRAD_TOL = 1e-3        # relative L2


def rel_l2(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    den = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / den) if den > 0 else float(np.linalg.norm(a))


def case_inputs(vol_shape, out_shape, R, seed=0, tf_name="rand", views=1, jitter=True):
    """Small seeded case: volume (1,D,H,W), tf (4,R), cameras (views,3), jitter (views,H,W) or None."""
    vol = make_volume(vol_shape, seed=1234 + seed)
    g = torch.Generator().manual_seed(100 + seed)
    if tf_name == "rand":
        tf = torch.rand(4, R, generator=g)
        tf[3] *= 0.15
    else:
        tf = make_tf(tf_name, R)
    cams = make_cameras(views, phase=0.3 + 0.1 * seed)
    w, h = out_shape
    jit = make_jitter(views, h, w, seed=4321 + seed) if jitter else None
    return vol, tf, cams, jit


def oracle_forward_views(vol, tf, cams, out_shape, jit=None, **kw):
    from oracle import cpu_oracle as co
    imgs, Ks, ns = [], [], []
    for v in range(cams.shape[0]):
        img, K, n = co.forward(vol.numpy(), tf.numpy(), cams[v].numpy(), out_shape,
                               jitter=None if jit is None else jit[v].numpy(), return_counts=True, **kw)
        imgs.append(img); Ks.append(K); ns.append(n)
    return np.stack(imgs), np.stack(Ks), np.stack(ns)


def oracle_backward_views(vol, tf, cams, grad_images, out_shape, jit=None, **kw):
    """Sum over views of the oracle's gradients (shared volume and TF)."""
    from oracle import cpu_oracle as co
    gv = gt = None
    for v in range(cams.shape[0]):
        a, b = co.backward(vol.numpy(), tf.numpy(), cams[v].numpy(), grad_images[v], out_shape,
                           jitter=None if jit is None else jit[v].numpy(), **kw)
        gv = a if gv is None else gv + a
        gt = b if gt is None else gt + b
    return gv, gt
