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
