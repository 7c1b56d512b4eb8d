"""Parity report: CUDA path (through the C ABI) vs the strict-fp32 CPU oracle on identical inputs and jitter, at sizes up to
the full C3 volume (layout "auto" = the cell-major copy with the exact empty-space skip grid, i.e. the default product path).
Prints one row per case: RGBA max-abs error, rays whose n / K differ, relative L2 of both gradients.
Run on a B200:  python tools/parity_report.py > gpurun_out/parity_report.txt"""
import os, sys, time
sys.path[:0] = [os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests")]
import numpy as np
import torch
from differender_b200 import VolumeRaycaster
from differender_b200.synthetic import make_cameras, make_jitter, make_tf, make_volume
from oracle import cpu_oracle as co

dev = "cuda:0"
rel = lambda a, b: float(np.linalg.norm(a.astype(np.float64) - b) / np.linalg.norm(b))
print(f"{'case':58s} {'rays':>8s} {'samples':>10s} {'rgba max|d|':>12s} {'alpha max|d|':>12s} {'K!=':>4s} {'gvol relL2':>11s} {'gtf relL2':>10s}")
CASES = [
    ("64^3 fp32, 96x64, tf1 R=128, sr 1, jitter, linear", 64, (96, 64), "tf1", 128, 1.0, True, torch.float32, "linear"),
    ("64^3 fp32, 96x64, rand TF R=33, sr 0.7, jitter", 64, (96, 64), "rand", 33, 0.7, True, torch.float32, "auto"),
    ("128^3 fp32, 128x128, tf1, sr 2, jitter", 128, (128, 128), "tf1", 128, 2.0, True, torch.float32, "auto"),
    ("128^3 fp16, 128x128, tf1, sr 1, jitter", 128, (128, 128), "tf1", 128, 1.0, True, torch.float16, "auto"),
    ("128^3 fp16, 128x128, tf1, sr 1, jitter, brick8", 128, (128, 128), "tf1", 128, 1.0, True, torch.float16, "brick8"),
    ("256^3 fp32 (C3 volume), 256x256, tf1, sr 1, jitter", 256, (256, 256), "tf1", 128, 1.0, True, torch.float32, "auto"),
    ("256^3 fp32 (C3 volume), 256x256, tf5, sr 1, no jitter", 256, (256, 256), "tf5", 128, 1.0, False, torch.float32, "auto"),
    ("256^3 fp32, 192x192, 'gray' TF (no transparent bins), sr 1", 256, (192, 192), "gray", 128, 1.0, True, torch.float32, "auto"),
    ("128^3 uint8 (DR_VOX_U8, 8-byte records), 128x128, tf1, sr 1", 128, (128, 128), "tf1", 128, 1.0, True, torch.uint8, "auto"),
    ("1100x24x24 fp16 (two-neighbour regime, direct taps), 64x48, tf1", (1100, 24, 24), (64, 48), "tf1", 128, 1.0, True, torch.float16, "auto"),
    ("24x24x1100 fp32 (two-neighbour regime, direct taps), 64x48, rand", (24, 24, 1100), (64, 48), "rand", 64, 1.0, True, torch.float32, "auto"),
]
for name, n, (w, h), tfn, R, sr, jitter, dt, layout in CASES:
    vol = make_volume(n)
    D, Hh, Ww = (n, n, n) if isinstance(n, int) else n
    if dt == torch.float16:
        vol = vol.half().float()
    if dt == torch.uint8:                                     # the oracle marches u8 / 255 in fp32 (numpy's float32 division)
        q = (vol * 255.0 + 0.5).to(torch.uint8)
        vol = torch.from_numpy((q.numpy().astype(np.float32) / np.float32(255.0)).astype(np.float32))
    tf = make_tf(tfn, R) if tfn != "rand" else torch.rand(4, R, generator=torch.Generator().manual_seed(1)) * torch.tensor([1, 1, 1, 0.15]).view(4, 1)
    cam = make_cameras(16)[3:4]
    jit = make_jitter(1, h, w) if jitter else None
    M = 8192
    vr = VolumeRaycaster((Ww, D, Hh), (w, h), max_samples=M, tf_resolution=R, layout=layout)
    v = vr.brick((q if dt == torch.uint8 else vol.to(dt)).to(dev).reshape(1, D, Hh, Ww).contiguous())
    tf_r4 = tf.to(dev).t().contiguous()[None]
    j = None if jit is None else jit.to(dev)
    out, K, Tp = vr.march(v, tf_r4, cam.to(dev), sr, j)
    go = torch.randn(out.shape, generator=torch.Generator().manual_seed(5))
    gv, gt = vr.march_backward(v, tf_r4, cam.to(dev), sr, j, go.to(dev), out, K, Tp, True, True)
    J = None if jit is None else jit[0].numpy()
    ref, Kr, nr = co.forward(vol.numpy(), tf.numpy(), cam[0].numpy(), (w, h), sampling_rate=sr, max_samples=M, jitter=J, return_counts=True)
    gvr, gtr = co.backward(vol.numpy(), tf.numpy(), cam[0].numpy(), go[0].numpy(), (w, h), sampling_rate=sr, max_samples=M, jitter=J)
    o = out[0].cpu().numpy(); Kc = K[0].cpu().numpy()
    same = Kc == Kr
    d = np.abs(o - ref)
    print(f"{name:58s} {w*h:8d} {int(Kr.sum()):10d} {d[:, same].max():12.2e} {d[3][same].max():12.2e} {int((~same).sum()):4d} "
          f"{rel(gv[0].cpu().numpy(), gvr):11.2e} {rel(gt[0].cpu().numpy().T, gtr):10.2e}", flush=True)
print("tolerances (BASELINE.json north_star): RGBA <= 1e-4 max-abs, gradients <= 1e-3 relative L2")

# --- the same CUDA path against what the REFERENCE'S OWN SOURCE computes on the fp32 interpreter (oracle/ti_shim.py; fixtures made by
#     tests/golden/make_shim_golden.py in the development container, where the reference tree is mounted)
import glob
print()
print("CUDA path vs the reference source executed on oracle/ti_shim.py (strict IEEE fp32, source order; tests/golden/shim/*.npz):")
print(f"{'case':58s} {'rays':>8s} {'samples':>10s} {'rgba max|d|':>12s} {'alpha max|d|':>12s} {'K!=':>4s} {'gvol relL2':>11s} {'gtf relL2':>10s}")
for f in sorted(glob.glob(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "shim", "*.npz"))):
    z = np.load(f)
    w, h = (int(v) for v in z["output_shape"])
    sr, M, nd = float(z["sampling_rate"]), int(z["max_samples"]), bool(z["nondiff"])
    D, Hh, Ww = z["volume"].shape
    vr = VolumeRaycaster((Ww, D, Hh), (w, h), max_samples=M, tf_resolution=z["tf"].shape[1], layout="auto", fov=float(z["fov"]), nearfar=(float(z["near"]), 100.0))
    v = vr.brick(torch.tensor(z["volume"]).to(dev).reshape(1, D, Hh, Ww).contiguous())
    tf_r4 = torch.tensor(z["tf"]).to(dev).t().contiguous()[None]
    cam = torch.tensor(z["cam"])[None].to(dev)
    j = torch.tensor(z["jitter"])[None].to(dev) if "jitter" in z.files else None
    out, K, Tp = vr.march(v, tf_r4, cam, sr, j, nondiff=nd)
    live = z["n"] > 1
    d = np.abs(out[0].cpu().numpy() - z["image"])[:, live]
    if nd:
        print(f"{str(z['name']):58s} {w*h:8d} {'':>10s} {d.max():12.2e} {d[3].max():12.2e}")
        continue
    go = torch.tensor(z["grad_image"])[None].to(dev)
    gv, gt = vr.march_backward(v, tf_r4, cam, sr, j, go, out, K, Tp, True, True)
    m = lambda a, b, nan: rel(np.where(nan, 0, a), np.where(nan, 0, b))
    print(f"{str(z['name']):58s} {w*h:8d} {int(z['K'].sum()):10d} {d.max():12.2e} {d[3].max():12.2e} {int((K[0].cpu().numpy() != z['K'])[live].sum()):4d} "
          f"{m(gv[0].cpu().numpy(), z['grad_volume'], z['gvol_nan']):11.2e} {m(gt[0].cpu().numpy().T, z['grad_tf'], z['gtf_nan']):10.2e}", flush=True)
