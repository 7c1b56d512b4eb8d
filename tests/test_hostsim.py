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


@pytest.mark.parametrize("shape,tf_name,sr,nondiff", [((64, 64, 64), "tf1", 1.0, False), ((40, 56, 72), "tf5", 1.0, False),
                                                     ((64, 64, 64), "tf1", 0.7, False), ((48, 48, 48), "tf1", 4.0, True),
                                                     ((1100, 9, 9), "tf1", 1.0, False)])
@pytest.mark.parametrize("layout", ["linear", "cell8"])
def test_empty_space_skipping_is_exact(shape, tf_name, sr, nondiff, layout):
    # transparent macro-cells are skipped as whole runs of samples: image, K and Tprev must not change by one bit
    out_shape = (40, 32)
    vol, tf, cams, jit = case_inputs(shape, out_shape, 128, seed=21, tf_name=tf_name, jitter=True)
    J = jit[0].numpy()
    grid = hs.skip_grid(vol.numpy(), tf.numpy(), out_shape, sampling_rate=sr, max_samples=4096)
    assert 0.02 < grid.mean() < 1.0                        # the case does have empty macro-cells, and non-empty ones
    kw = dict(sampling_rate=sr, max_samples=4096, jitter=J, nondiff=nondiff, cell=layout == "cell8")
    out, K, Tp, n = hs.forward(vol.numpy(), tf.numpy(), cams[0].numpy(), out_shape, **kw)
    out2, K2, Tp2, n2 = hs.forward(vol.numpy(), tf.numpy(), cams[0].numpy(), out_shape, skip=True, **kw)
    assert np.array_equal(out, out2) and np.array_equal(K, K2) and np.array_equal(Tp, Tp2) and np.array_equal(n, n2)
    ref, Kr, _ = co.forward(vol.numpy(), tf.numpy(), cams[0].numpy(), out_shape, sampling_rate=sr, max_samples=4096, jitter=J, nondiff=nondiff,
                            return_counts=True)
    assert np.array_equal(K2, Kr) and np.array_equal(ref[3], out2[3])


def test_skip_grid_never_marks_a_cell_that_can_be_opaque():
    # brute force: every voxel whose TF bins (its own and the next) have non-zero alpha must lie in a non-empty macro-cell
    vol, tf, _, _ = case_inputs((40, 56, 72), (8, 8), 64, seed=4, tf_name="tf3", jitter=False)
    g = hs.skip_grid(vol.numpy(), tf.numpy(), (8, 8))
    v = vol.numpy().reshape(40, 56, 72)                    # [y][z][x]
    a = tf.numpy()[3]
    lo = np.minimum(np.floor(np.maximum(v, 0) * 63).astype(int), 63)
    opaque = (a[lo] != 0) | (a[np.minimum(lo + 1, 63)] != 0)
    ys, zs, xs = np.nonzero(opaque)
    E = 40 // g.shape[0] if 40 % g.shape[0] == 0 else 8     # macro-cell edge (8 cells; 4 in the DR_MACRO_SHIFT=2 variant)
    for dy in (0, -1):                                     # a voxel is a corner of cells in its own and the previous macro-cell row
        for dz in (0, -1):
            for dx in (0, -1):
                assert not g[np.maximum(ys + dy, 0) // E, np.maximum(zs + dz, 0) // E, np.maximum(xs + dx, 0) // E].any()


def test_empty_space_skipping_respects_max_samples_and_views_without_hits():
    # max_samples cuts the rays inside an empty run; a camera far to the side leaves most rays without samples
    vol, tf, cams, jit = case_inputs((64, 64, 64), (40, 32), 128, seed=2, tf_name="tf1", jitter=True)
    J = jit[0].numpy()
    for M in (1, 7, 33):
        a = hs.forward(vol.numpy(), tf.numpy(), cams[0].numpy(), (40, 32), max_samples=M, jitter=J)
        b = hs.forward(vol.numpy(), tf.numpy(), cams[0].numpy(), (40, 32), max_samples=M, jitter=J, skip=True)
        assert all(np.array_equal(x, y) for x, y in zip(a, b)) and a[1].max() == M
    cam = np.array([9.0, 0.3, 0.2], np.float32)
    a = hs.forward(vol.numpy(), tf.numpy(), cam, (40, 32), max_samples=2048, jitter=J)
    b = hs.forward(vol.numpy(), tf.numpy(), cam, (40, 32), max_samples=2048, jitter=J, skip=True)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))


@pytest.mark.parametrize("shape,tf_name,sr", [((64, 64, 64), "tf1", 1.0), ((40, 56, 72), "tf5", 1.0), ((64, 64, 64), "tf1", 0.7), ((1100, 9, 9), "tf1", 1.0)])
@pytest.mark.parametrize("layout", ["linear", "cell8"])
def test_volume_only_backward_with_skip_grid_is_exact(shape, tf_name, sr, layout):
    # the volume-only backward jumps over runs of samples in macro-cells that are transparent under the TF: they contribute
    # nothing, so the cell-major gradient must be IDENTICAL (same additions in the same order on the host) with and without the grid
    out_shape = (40, 32)
    vol, tf, cams, jit = case_inputs(shape, out_shape, 128, seed=21, tf_name=tf_name, jitter=True)
    J = jit[0].numpy()
    go = np.random.default_rng(3).normal(size=(4, 32, 40)).astype(np.float32)
    kw = dict(sampling_rate=sr, max_samples=4096, jitter=J, want_vol=True, want_tf=False, cell=layout == "cell8")
    gv, gt = hs.backward(vol.numpy(), tf.numpy(), cams[0].numpy(), go, out_shape, **kw)
    gv2, gt2 = hs.backward(vol.numpy(), tf.numpy(), cams[0].numpy(), go, out_shape, skip=True, **kw)
    assert np.abs(gv).max() > 0 and np.array_equal(gv, gv2) and not gt2.any()
    ref, _ = co.backward(vol.numpy(), tf.numpy(), cams[0].numpy(), go, out_shape, sampling_rate=sr, max_samples=4096, jitter=J, want_tf=False)
    assert rel_l2(gv2, ref) <= 1e-4
    # with a TF gradient the grid must be ignored (transparent samples feed d(alpha))
    a = hs.backward(vol.numpy(), tf.numpy(), cams[0].numpy(), go, out_shape, sampling_rate=sr, max_samples=4096, jitter=J, cell=layout == "cell8")
    b = hs.backward(vol.numpy(), tf.numpy(), cams[0].numpy(), go, out_shape, sampling_rate=sr, max_samples=4096, jitter=J, cell=layout == "cell8", skip=True)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
