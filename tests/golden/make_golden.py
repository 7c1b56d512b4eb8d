"""Generates tests/golden/*.npz: seeded inputs plus the ORACLE's outputs for them.

PARITY UNPINNED: the real reference (Taichi) cannot run here, so these are not outputs of the reference
itself.  Each fixture stores the strict-fp32 C oracle's image / K / gradients, after this script has checked them against
the independent float64 torch.autograd restatement (oracle/torch_ref.py) -- the check that pins the oracle.
Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

from helpers import case_inputs, rel_l2  # noqa: E402
from oracle import cpu_oracle as co, torch_ref as tr  # noqa: E402

CASES = {
    "g0_32c_48x40_r16_jit": dict(shape=(32, 32, 32), out=(48, 40), R=16, views=2, jitter=True, sr=1.0, M=2048, tf="rand"),
    "g1_ragged_r33_nojit": dict(shape=(37, 29, 45), out=(50, 34), R=33, views=1, jitter=False, sr=1.0, M=2048, tf="rand"),
    "g2_48c_tf1_sr07": dict(shape=(48, 48, 48), out=(40, 40), R=128, views=1, jitter=True, sr=0.7, M=2048, tf="tf1"),
}

for name, c in CASES.items():
    vol, tf, cams, jit = case_inputs(c["shape"], c["out"], c["R"], seed=len(name), tf_name=c["tf"], views=c["views"], jitter=c["jitter"])
    g = torch.Generator().manual_seed(11)
    imgs, Ks, gos = [], [], []
    gv = gt = None
    for v in range(c["views"]):
        J = None if jit is None else jit[v]
        kw = dict(sampling_rate=c["sr"], max_samples=c["M"])
        img, K, n = co.forward(vol.numpy(), tf.numpy(), cams[v].numpy(), c["out"], jitter=None if J is None else J.numpy(), return_counts=True, **kw)
        go = torch.randn(img.shape, generator=g)
        a, b = co.backward(vol.numpy(), tf.numpy(), cams[v].numpy(), go.numpy(), c["out"], jitter=None if J is None else J.numpy(), **kw)
        # pin against the independent fp64 autograd restatement
        V = vol.clone().double().requires_grad_(True); T = tf.clone().double().requires_grad_(True)
        im64, K64, n64 = tr.render(V, T, cams[v], c["out"], jitter=J, dtype=torch.float64, return_counts=True, **kw)
        (im64 * go.double()).sum().backward()
        same = (K64.numpy() == K) & (n64.numpy() == n)       # H6: discrete decisions can flip between fp32 and fp64
        err = np.abs(im64.detach().numpy() - img)[:, same].max()
        a64, b64 = co.backward(vol.numpy(), tf.numpy(), cams[v].numpy(), go.numpy(), c["out"], jitter=None if J is None else J.numpy(), fp64=True, **kw)
        print(f"{name} view {v}: K/n mismatches vs fp64 {int((~same).sum())}, image err vs fp64 {err:.2e}, "
              f"fp64-C vs autograd: vol {rel_l2(a64, V.grad.numpy()):.1e} tf {rel_l2(b64, T.grad.numpy()):.1e}; "
              f"fp32-C vs autograd: vol {rel_l2(a, V.grad.numpy()):.1e} tf {rel_l2(b, T.grad.numpy()):.1e}")
        img64 = co.forward(vol.numpy(), tf.numpy(), cams[v].numpy(), c["out"], jitter=None if J is None else J.numpy(), fp64=True, **kw)
        pin = np.abs(img64 - im64.detach().numpy()).max()
        # the formulas are pinned in float64 (same C source as the fp32 oracle); fp32-vs-fp64 differences are the
        # conditioning of the +-1e-3 central-difference normal, reported and loosely bounded
        assert pin < 1e-9 and rel_l2(a64, V.grad.numpy()) < 1e-6 and rel_l2(b64, T.grad.numpy()) < 1e-6, (pin,)
        assert err < 5e-3
        imgs.append(img); Ks.append(K); gos.append(go.numpy())
        gv = a if gv is None else gv + a
        gt = b if gt is None else gt + b
    out = dict(volume=vol.numpy(), tf=tf.numpy(), cams=cams.numpy(), output_shape=np.array(c["out"]), sampling_rate=np.float64(c["sr"]),
               max_samples=np.int64(c["M"]), image=np.stack(imgs), K=np.stack(Ks), grad_image=np.stack(gos),
               grad_volume=gv.astype(np.float32), grad_tf=gt.astype(np.float32))
    if jit is not None:
        out["jitter"] = jit.numpy()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), **out)
print("ok")
