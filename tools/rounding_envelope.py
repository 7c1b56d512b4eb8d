#!/usr/bin/env python
"""Rounding envelope of the oracle.

The reference's arithmetic is Taichi's (fast_math=True, LLVM contraction, libdevice fast pow/tan), which cannot run here, so
oracle/cpu_ref.c pins ONE rounding per open choice.  This script builds the oracle once per alternative choice
(oracle/cpu_oracle.py VARIANTS -> -DORA_* switches of cpu_ref.c) and reports how far each alternative moves the image and
the gradients from the default build on C3-like inputs (synthetic volume seed 1234, tf1, in_circles cameras, supplied
jitter, sampling rate 1) -- i.e. how much of the north_star tolerance (RGBA 1e-4 max-abs, gradients 1e-3 relative L2)
the unresolved choices can consume.  CPU only; test infrastructure (reads oracle/, never the product).

    python tools/rounding_envelope.py [--out profiles/r02_rounding_envelope.txt] [--quick]
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from differender_b200.synthetic import make_cameras, make_jitter, make_tf, make_volume  # noqa: E402
from oracle import cpu_oracle as co  # noqa: E402


def rel_l2(a, b):
    den = np.linalg.norm(b)
    return float(np.linalg.norm(np.asarray(a, np.float64) - b) / den) if den > 0 else float(np.linalg.norm(a))


def run(variant, vol, tf, cams, jit, res, M, fp64=False):
    imgs, Ks, ns, gv, gt = [], [], [], None, None
    for v in range(cams.shape[0]):
        kw = dict(max_samples=M, jitter=jit[v], variant=variant, fp64=fp64)
        img, K, n = co.forward(vol, tf, cams[v], res, return_counts=True, **kw)
        go = np.random.default_rng(7 + v).standard_normal(img.shape).astype(np.float32)
        a, b = co.backward(vol, tf, cams[v], go, res, **kw)
        imgs.append(img); Ks.append(K); ns.append(n)
        gv = a if gv is None else gv + a
        gt = b if gt is None else gt + b
    return np.stack(imgs).astype(np.float64), np.stack(Ks), np.stack(ns), gv, gt


def compare(name, ref, got):
    (i0, K0, n0, gv0, gt0), (i1, K1, n1, gv1, gt1) = ref, got
    same = K0 == K1
    d = np.abs(i1 - i0).max(axis=1)                                  # per pixel, over channels
    return dict(name=name, n_diff=int((n0 != n1).sum()), K_diff=int((~same).sum()), rays=int(same.size),
                max_abs=float(d[same].max()), max_abs_all=float(d.max()), within=float((d <= 1e-4).mean()),
                p999=float(np.quantile(d, 0.999)), gvol=rel_l2(gv1, gv0), gtf=rel_l2(gt1, gt0))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r02_rounding_envelope.txt"))
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    co.build()
    cases = [("C3-like: 256^3 fp32, tf1, 256x256 rays x 2 views", 256, "tf1", (256, 256), 2, 2048),
             ("dense TF (every sample shaded): 256^3, gray, 192x192 rays x 1 view", 256, "gray", (192, 192), 1, 2048),
             ("small volume (normal offset 0.03 voxel: worst cancellation): 64^3, tf1, 256x256 rays x 2 views", 64, "tf1", (256, 256), 2, 2048)]
    if args.quick:
        cases = [("quick: 64^3, tf1, 96x96 x 1", 64, "tf1", (96, 96), 1, 2048)]
    lines = ["# Rounding envelope of oracle/cpu_ref.c  (tools/rounding_envelope.py; CPU, strict fp32 builds, -ffp-contract=off)",
             "# each row: one alternative rounding choice (cpu_ref.c header) vs the DEFAULT oracle build on identical inputs and jitter.",
             "# n_diff / K_diff: rays whose sample count n / active-sample count K differ (discrete decisions, SURVEY H6);",
             "# max_abs: max |RGBA difference| over rays with equal K (all: over all rays); within: fraction of rays whose RGBA",
             "# differs by <= 1e-4; p99.9: 99.9th percentile of the per-ray difference; gvol / gtf: relative L2 of the volume / TF",
             "# gradient vs the default build.  north_star tolerances: RGBA 1e-4 max-abs, gradients 1e-3 relative L2.",
             "# `fp64` = the same source in double (the 'true' value both fp32 builds approximate).", ""]
    worst = dict(max_abs=0.0, within=1.0, gvol=0.0, gtf=0.0)
    for title, n, tfname, res, views, M in cases:
        t0 = time.time()
        vol = make_volume(n).numpy()[0]
        tf = make_tf(tfname, 128).numpy()
        cams = make_cameras(16)[:views].numpy()
        jit = make_jitter(views, res[1], res[0]).numpy()
        ref = run(None, vol, tf, cams, jit, res, M)
        rows = [compare("fp64", ref, run(None, vol, tf, cams, jit, res, M, fp64=True))]
        for name in co.VARIANTS:
            rows.append(compare(name, ref, run(name, vol, tf, cams, jit, res, M)))
        lines.append(f"## {title}   ({int(ref[1].sum())} active samples, {time.time() - t0:.0f} s)")
        lines.append(f"{'variant':<20}{'n_diff':>8}{'K_diff':>8}{'max_abs':>11}{'max_abs(all)':>14}{'within 1e-4':>13}{'p99.9':>11}{'gvol relL2':>12}{'gtf relL2':>12}")
        for r in rows:
            lines.append(f"{r['name']:<20}{r['n_diff']:>8}{r['K_diff']:>8}{r['max_abs']:>11.2e}{r['max_abs_all']:>14.2e}{100 * r['within']:>12.3f}%"
                         f"{r['p999']:>11.2e}{r['gvol']:>12.2e}{r['gtf']:>12.2e}")
            if r["name"] != "fp64":
                worst["max_abs"] = max(worst["max_abs"], r["max_abs_all"]); worst["within"] = min(worst["within"], r["within"])
                worst["gvol"] = max(worst["gvol"], r["gvol"]); worst["gtf"] = max(worst["gtf"], r["gtf"])
        lines.append("")
        print("\n".join(lines[-(len(rows) + 3):]), flush=True)
    lines.append(f"ENVELOPE (worst over all variants and cases, fp64 row excluded): max-abs RGBA {worst['max_abs']:.2e}, "
                 f"{100 * worst['within']:.3f} % of rays within 1e-4, volume-gradient rel-L2 {worst['gvol']:.2e}, TF-gradient rel-L2 {worst['gtf']:.2e}")
    print(lines[-1])
    if not args.quick:
        with open(args.out, "w") as f:
            f.write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
