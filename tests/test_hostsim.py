"""Device arithmetic (differender_b200/csrc/dr_math.cuh, compiled for the host by tests/hostsim) against the oracle.
Checks what the GPU tests cannot attribute: the corner-reuse taps are bit-identical to full trilinear taps, the tape-free
reverse march equals the taped adjoint, the merged 32-voxel scatter equals the 56-weight scatter."""
import numpy as np
import pytest

import hostsim_lib as hs
from helpers import case_inputs, rel_l2
from oracle import cpu_oracle as co


@pytest.mark.parametrize("shape,out_shape,R,sr,jitter", [
    ((32, 32, 32), (48, 40), 16, 1.0, True),
    ((37, 29, 45), (33, 29), 128, 1.0, False),
    ((24, 24, 24), (32, 32), 32, 0.7, True),
])
@pytest.mark.parametrize("generic,brick,cell", [(False, False, False), (True, False, False), (False, True, False), (False, False, True)])
def test_device_math_matches_oracle(shape, out_shape, R, sr, jitter, generic, brick, cell):
    vol, tf, cams, jit = case_inputs(shape, out_shape, R, seed=R, jitter=jitter)
    J = None if jit is None else jit[0].numpy()
    img, K, n = co.forward(vol.numpy(), tf.numpy(), cams[0].numpy(), out_shape, sampling_rate=sr, jitter=J, return_counts=True)
    out, K2, Tp, n2 = hs.forward(vol.numpy(), tf.numpy(), cams[0].numpy(), out_shape, sampling_rate=sr, jitter=J, generic=generic, brick=brick, cell=cell)
    assert np.array_equal(n, n2) and np.array_equal(K, K2)
    assert np.array_equal(img[3], out[3])                      # alpha path: bit-identical by construction
    assert np.abs(img - out).max() <= 1e-6
    go = np.random.default_rng(5).normal(size=img.shape).astype(np.float32)
    gv, gt = co.backward(vol.numpy(), tf.numpy(), cams[0].numpy(), go, out_shape, sampling_rate=sr, jitter=J)
    gv2, gt2 = hs.backward(vol.numpy(), tf.numpy(), cams[0].numpy(), go, out_shape, sampling_rate=sr, jitter=J, generic=generic, brick=brick, cell=cell)
    assert rel_l2(gv2, gv) <= 1e-4 and rel_l2(gt2, gt) <= 1e-4


@pytest.mark.parametrize("shape", [(1100, 6, 6), (6, 1100, 6), (6, 6, 1100)])
@pytest.mark.parametrize("layout", ["linear", "brick8", "cell8"])
def test_device_math_both_taps_of_an_axis_cross(shape, layout):
    brick, cell = layout == "brick8", layout == "cell8"
    # an axis longer than 1000 voxels: the +-1e-3 taps move more than half a voxel, so BOTH can leave the centre cell
    # (the branch behind the select-based tap evaluation); still below the generic-tap threshold (~2000)
    out_shape = (20, 16)
    vol, tf, cams, jit = case_inputs(shape, out_shape, 32, seed=3, jitter=True)
    J = jit[0].numpy()
    img, K, n = co.forward(vol.numpy(), tf.numpy(), cams[0].numpy(), out_shape, jitter=J, max_samples=4096, return_counts=True)
    out, K2, Tp, n2 = hs.forward(vol.numpy(), tf.numpy(), cams[0].numpy(), out_shape, jitter=J, max_samples=4096, brick=brick, cell=cell)
    assert n.max() > 1000 and np.array_equal(n, n2) and np.array_equal(K, K2)
    assert np.array_equal(img[3], out[3]) and np.abs(img - out).max() <= 1e-6
    go = np.random.default_rng(9).normal(size=img.shape).astype(np.float32)
    gv, gt = co.backward(vol.numpy(), tf.numpy(), cams[0].numpy(), go, out_shape, jitter=J, max_samples=4096)
    gv2, gt2 = hs.backward(vol.numpy(), tf.numpy(), cams[0].numpy(), go, out_shape, jitter=J, max_samples=4096, brick=brick, cell=cell)
    assert rel_l2(gv2, gv) <= 1e-4 and rel_l2(gt2, gt) <= 1e-4


def test_device_math_nondiff_and_truncation():
    vol, tf, cams, _ = case_inputs((32, 32, 32), (40, 40), 64, seed=8, tf_name="tf1", jitter=False)
    img, K, n = co.forward(vol.numpy(), tf.numpy(), cams[0].numpy(), (40, 40), sampling_rate=4.0, nondiff=True, return_counts=True)
    out, K2, _, n2 = hs.forward(vol.numpy(), tf.numpy(), cams[0].numpy(), (40, 40), sampling_rate=4.0, nondiff=True)
    assert np.array_equal(K, K2) and np.array_equal(n, n2) and np.abs(img - out).max() <= 1e-6
    img, K, n = co.forward(vol.numpy(), tf.numpy(), cams[0].numpy(), (40, 40), max_samples=17, return_counts=True)
    out, K2, _, _ = hs.forward(vol.numpy(), tf.numpy(), cams[0].numpy(), (40, 40), max_samples=17)
    assert K.max() == 17 and np.array_equal(K, K2) and np.abs(img - out).max() <= 1e-6


@pytest.mark.parametrize("want_vol,want_tf", [(True, False), (False, True)])
def test_device_math_single_gradient_with_transparent_tf(want_vol, want_tf):
    # tf1 is exactly transparent between its bumps: exercises the transparent-sample shortcuts of both marches
    vol, tf, cams, jit = case_inputs((32, 32, 32), (40, 32), 128, seed=12, tf_name="tf1", jitter=True)
    J = jit[0].numpy()
    img, K, n = co.forward(vol.numpy(), tf.numpy(), cams[0].numpy(), (40, 32), jitter=J, max_samples=2048, return_counts=True)
    out, K2, Tp, n2 = hs.forward(vol.numpy(), tf.numpy(), cams[0].numpy(), (40, 32), jitter=J, max_samples=2048)
    assert np.array_equal(K, K2) and np.array_equal(img[3], out[3]) and np.abs(img - out).max() <= 1e-6
    go = np.random.default_rng(6).normal(size=img.shape).astype(np.float32)
    gv, gt = co.backward(vol.numpy(), tf.numpy(), cams[0].numpy(), go, (40, 32), jitter=J, max_samples=2048)
    gv2, gt2 = hs.backward(vol.numpy(), tf.numpy(), cams[0].numpy(), go, (40, 32), jitter=J, max_samples=2048,
                           want_vol=want_vol, want_tf=want_tf)
    if want_vol:
        assert rel_l2(gv2, gv) <= 1e-4 and not gt2.any()
    else:
        assert rel_l2(gt2, gt) <= 1e-4 and not gv2.any()
