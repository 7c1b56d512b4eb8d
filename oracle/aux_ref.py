"""numpy restatements of the caller-side steps next to the march (TEST INFRASTRUCTURE, not product code).

momentum_step: reference examples/taichi_volume_raycaster.py:375-381 (`apply_grad`), float32, one rounding per operator.
ingest_u8:     reference examples/taichi_volume_raycaster.py:548-550 (skull.raw -> swapaxes(0,1) -> /255).
mse:           torch.nn.functional.mse_loss as used at :432-435 and examples/test_opt_tf.py:70-72.
"""
import numpy as np


def momentum_step(param, grad, momentum, lr, gamma, max_grad, lo=0.0, hi=np.inf):
    f = np.float32
    g = np.clip(grad.astype(f), f(-max_grad), f(max_grad))                          # tl.clamp(grad, -max_grad, max_grad)
    m = (f(gamma) * momentum.astype(f)).astype(f) + (f(lr) * g).astype(f)            # gamma * m + lr * clamp(...)
    m = m.astype(f)
    p = np.clip((param.astype(f) - m).astype(f), f(lo), f(hi) if np.isfinite(hi) else None)   # tf -= m ; tf = max(tf, 0)
    return p.astype(f), m


def ingest_u8(raw, swap_axes01=True):
    v = np.swapaxes(raw, 0, 1) if swap_axes01 else raw
    return (v.astype(np.float32) / np.float32(255.0)).astype(np.float32)


def mse(img, target):
    d = img.astype(np.float64) - target.astype(np.float64)
    return float((d * d).mean()), (2.0 * d / d.size)


def ssim(X, Y, data_range=1.0, win_size=11, win_sigma=1.5, K=(0.01, 0.03), nonnegative=False):
    """SSIM of two (N, C, H, W) batches as pytorch_msssim.ssim defines it (reference examples/test_opt_tf.py:70), in float64
    with explicit loops over the window: 'valid' separable Gaussian filtering, per-channel mean of the SSIM map, optional ReLU,
    mean over channels and images."""
    X = np.asarray(X, np.float64); Y = np.asarray(Y, np.float64)
    g = np.exp(-((np.arange(win_size) - win_size // 2) ** 2) / (2.0 * win_sigma ** 2)); g /= g.sum()

    def blur(a):
        if a.shape[2] >= win_size:
            a = sum(g[k] * a[:, :, k:a.shape[2] - win_size + 1 + k, :] for k in range(win_size))
        if a.shape[3] >= win_size:
            a = sum(g[k] * a[:, :, :, k:a.shape[3] - win_size + 1 + k] for k in range(win_size))
        return a
    C1, C2 = (K[0] * data_range) ** 2, (K[1] * data_range) ** 2
    m1, m2 = blur(X), blur(Y)
    s1, s2, s12 = blur(X * X) - m1 * m1, blur(Y * Y) - m2 * m2, blur(X * Y) - m1 * m2
    smap = ((2 * m1 * m2 + C1) / (m1 * m1 + m2 * m2 + C1)) * ((2 * s12 + C2) / (s1 + s2 + C2))
    pc = smap.reshape(smap.shape[0], smap.shape[1], -1).mean(-1)
    if nonnegative:
        pc = np.maximum(pc, 0.0)
    return float(pc.mean())
