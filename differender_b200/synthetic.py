"""Seeded synthetic inputs for the benchmark configs and the parity tests.

The reference's bundled volumes (examples/data/skull.raw, skewed_head.*) are missing from the mount
(.MISSING_LARGE_BLOBS), so every config uses this stand-in: a smooth field of sine products plus Gaussian
blobs plus small uniform noise (no exactly-flat cells: the central-difference normal is 0/0 there).
"""
import math

import torch

from .utils import get_tf, in_circles

__all__ = ["make_volume", "make_cameras", "make_jitter", "make_tf"]


def make_volume(n, seed=1234, device="cpu", dtype=torch.float32, noise=1e-2, chunk=64):
    """(1, D, H, W) volume with D=H=W=n (or n = (D,H,W)), values in (0,1).  Deterministic per (n, seed, device type)."""
    D, H, W = (n, n, n) if isinstance(n, int) else n
    g = torch.Generator(device="cpu").manual_seed(seed)
    freqs = torch.rand(3, 3, generator=g) * 2.5 + 0.75          # cycles over [-1,1]
    phases = torch.rand(3, 3, generator=g) * 2 * math.pi
    centers = torch.rand(8, 3, generator=g) * 1.4 - 0.7
    widths = torch.rand(8, generator=g) * 0.25 + 0.15
    amps = torch.rand(8, generator=g) * 0.5 + 0.25
    dev = torch.device(device)
    freqs, phases, centers, widths, amps = (t.to(dev) for t in (freqs, phases, centers, widths, amps))
    gn = torch.Generator(device=dev).manual_seed(seed + 1)
    out = torch.empty((1, D, H, W), dtype=dtype, device=dev)
    xs = torch.linspace(-1, 1, W, device=dev).view(1, 1, W)
    ys = torch.linspace(-1, 1, H, device=dev).view(1, H, 1)
    for z0 in range(0, D, chunk):                                # chunked along D so 1024^3 never needs > a few GiB of temporaries
        z1 = min(D, z0 + chunk)
        zs = torch.linspace(-1, 1, D, device=dev)[z0:z1].view(-1, 1, 1)
        v = torch.zeros((z1 - z0, H, W), device=dev)
        for k in range(3):
            v = v + torch.sin(math.pi * freqs[k, 0] * xs + phases[k, 0]) * torch.sin(math.pi * freqs[k, 1] * ys + phases[k, 1]) \
                * torch.sin(math.pi * freqs[k, 2] * zs + phases[k, 2])
        v = 0.32 + 0.08 * v
        for b in range(8):
            r2 = (xs - centers[b, 0]) ** 2 + (ys - centers[b, 1]) ** 2 + (zs - centers[b, 2]) ** 2
            v = v + 0.45 * amps[b] * torch.exp(-r2 / (2 * widths[b] ** 2))
        # fade to a low-intensity shell so the rays do not saturate at the box face
        rr = torch.sqrt(xs ** 2 + ys ** 2 + zs ** 2)
        v = v * torch.clamp(1.35 - rr, 0.0, 1.0)
        v = v + noise * torch.rand(v.shape, generator=gn, device=dev)
        out[0, z0:z1] = torch.clamp(v, 0.0, 1.0).to(dtype)
    return out


def make_tf(name="tf1", res=128, device="cpu"):
    """(4, res) preset transfer function (reference utils.py:9-65)."""
    return get_tf(name, res).to(device)


def make_cameras(bs, device="cpu", phase=0.0):
    """(bs, 3): view v = in_circles(2*pi*v/bs + phase) (reference utils.py:80-83)."""
    return torch.stack([in_circles(2 * math.pi * v / bs + phase) for v in range(bs)]).to(device)


def make_jitter(bs, h, w, seed=4321, device="cpu"):
    g = torch.Generator(device=torch.device(device)).manual_seed(seed)
    return torch.rand((bs, h, w), generator=g, device=device)
