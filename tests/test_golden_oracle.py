"""The strict-fp32 oracle must reproduce the committed golden fixtures bit for bit on images and K (they were produced by
it, after being pinned against the fp64 autograd restatement -- tests/golden/make_golden.py)."""
import glob
import os

import numpy as np

from helpers import rel_l2
from oracle import cpu_oracle as co


def test_oracle_reproduces_golden_fixtures():
    files = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))
    assert len(files) >= 3
    for f in files:
        z = np.load(f)
        out_shape = tuple(int(v) for v in z["output_shape"])
        kw = dict(sampling_rate=float(z["sampling_rate"]), max_samples=int(z["max_samples"]))
        gv = gt = None
        for v in range(z["cams"].shape[0]):
            J = z["jitter"][v] if "jitter" in z.files else None
            img, K, _ = co.forward(z["volume"], z["tf"], z["cams"][v], out_shape, jitter=J, return_counts=True, **kw)
            assert np.array_equal(K, z["K"][v]), f
            assert np.array_equal(img, z["image"][v]), f
            a, b = co.backward(z["volume"], z["tf"], z["cams"][v], z["grad_image"][v], out_shape, jitter=J, **kw)
            gv = a if gv is None else gv + a
            gt = b if gt is None else gt + b
        assert rel_l2(gv, z["grad_volume"]) < 1e-6 and rel_l2(gt, z["grad_tf"]) < 1e-6, f
