"""The opportunistic real-Taichi probe (oracle/taichi_probe.py) and the oracle's rounding variants: host-only checks.
Taichi is not installable in this image, so what can be verified here is that the probe degrades to a clear message, that its
three-line jitter patch still applies to the reference source (when /root/reference is mounted -- it is not on the GPU box),
and that every rounding switch of oracle/cpu_ref.c really selects different arithmetic while staying inside a sane envelope."""
import ast

import numpy as np
import pytest

from helpers import case_inputs
from oracle import cpu_oracle as co
from oracle import taichi_probe as tp


def test_probe_reports_unavailable_without_taichi():
    ok, msg = tp.taichi_status()
    st = tp.status()
    assert st["taichi"] == ok and isinstance(st["note"], str) and st["note"]
    if not ok:
        assert "Taichi unavailable" in msg and st["pinned"] is False and "unpinned" in st["note"]
        assert tp.main() == 0


def test_jitter_patch_applies_to_the_reference_source():
    ref = tp.find_reference()
    if ref is None:
        pytest.skip("reference source not mounted (GPU box)")
    src = tp.patched_source(ref)
    ast.parse(src)
    assert "ti.random" not in src and src.count("self.jitter_field") == 3


def test_rounding_variants_are_distinct_and_small():
    vol, tf, cams, jit = case_inputs((24, 24, 24), (32, 24), 32, seed=3, tf_name="rand", views=1)
    v, t, c, j = vol.numpy(), tf.numpy(), cams[0].numpy(), jit[0].numpy()
    ref = co.forward(v, t, c, (32, 24), jitter=j, max_samples=512, sampling_rate=0.7)
    go = np.ones_like(ref)
    gv0, gt0 = co.backward(v, t, c, go, (32, 24), jitter=j, max_samples=512, sampling_rate=0.7)
    for name in co.VARIANTS:
        img = co.forward(v, t, c, (32, 24), jitter=j, max_samples=512, sampling_rate=0.7, variant=name)
        gv, gt = co.backward(v, t, c, go, (32, 24), jitter=j, max_samples=512, sampling_rate=0.7, variant=name)
        d = float(np.abs(img - ref).max())
        changed = d > 0 or not np.array_equal(gv, gv0) or not np.array_equal(gt, gt0)
        assert changed, f"{name}: the switch does not change the arithmetic"
        assert d <= 5e-4, f"{name}: {d}"                     # an envelope, not a parity bar: rounding-level, not a different algorithm
