#!/usr/bin/env python
"""tests/golden/shim/*.npz: outputs of the REFERENCE'S OWN SOURCE (differender/volume_raycaster.py, read from /root/reference or
$DIFFERENDER_REFERENCE, with the probe's 3-line jitter patch) executed on oracle/ti_shim.py, the strict-IEEE-fp32 interpreter of
the Taichi subset it uses.  Run in the development container (the reference tree does not travel to the GPU box):

    python tests/golden/make_shim_golden.py

Each file holds the inputs (volume (D,H,W), tf (4,R), cam (3,), jitter (H,W) if any, grad_image (4,H,W) with the seeds of n <= 1 rays
zeroed), the parameters, and what the reference source computed: image (4,H,W), K, n (H,W), grad_volume (D,H,W) / grad_tf (4,R) in
float64 after its nan_to_num, and the masks of the entries it had NaN-poisoned before that.  Forward-only
(`nondiff`) cases hold image and n."""
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import taichi_probe as tp          # noqa: E402
from oracle import ti_shim                     # noqa: E402


def main():
    if tp.find_reference() is None:
        sys.exit("the reference source was not found (set DIFFERENDER_REFERENCE)")
    mod = tp._load_reference_module("shim")
    out_dir = os.path.join(HERE, "shim")
    os.makedirs(out_dir, exist_ok=True)
    for k, c in enumerate(tp.SHIM_CASES):
        vol, tf, cam, jit, go = tp.case_inputs(c)
        ti_shim.reset()
        nondiff = bool(c.get("nondiff"))
        r = tp.run_reference(mod, vol, tf, cam, None if nondiff else jit, None if nondiff else go, c["res"], c["M"], c["sr"], nondiff=nondiff)
        z = dict(name=c["name"], volume=vol, tf=tf, cam=cam, output_shape=np.array(c["res"]), sampling_rate=c["sr"], max_samples=c["M"],
                 nondiff=nondiff, image=r["image"], n=r["n"])
        if not nondiff:
            if jit is not None:
                z["jitter"] = jit
            z.update(grad_image=r["grad_image"], K=r["K"], grad_volume=r["gvol"], grad_tf=r["gtf"], gvol_nan=r["gvol_nan"], gtf_nan=r["gtf_nan"])
        slug = re.sub(r"[^a-z0-9]+", "_", c["name"].lower()).strip("_")
        path = os.path.join(out_dir, f"s{k}_{slug}.npz")
        np.savez_compressed(path, **z)
        print(f"{path}: {os.path.getsize(path)} bytes, rays with samples {int((r['n'] > 0).sum())}, longest {int(r['n'].max())}")


if __name__ == "__main__":
    main()
