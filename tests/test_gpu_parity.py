"""Parity of the CUDA path (through the C ABI, via the Python host) against the CPU oracle on identical inputs and jitter.
Tolerances are the north_star's: RGBA <= 1e-4 max-abs, TF / volume gradients <= 1e-3 relative L2."""
import numpy as np
import pytest
import torch

from helpers import GRAD_TOL, RGBA_TOL, case_inputs, oracle_backward_views, oracle_forward_views, rel_l2

pytestmark = pytest.mark.gpu


# The product default is layout="auto" (the cell-major copy, `cell8`, at these sizes) with the exact empty-space skip grid; the
# zero-copy `linear` layout is the variant.  Every oracle comparison below runs on the default unless it names a layout.
DEFAULT_LAYOUT = "auto"


def _vr(vol, out_shape, R, M, layout=DEFAULT_LAYOUT, fov=30.0, near=0.1):
    from differender_b200 import VolumeRaycaster
    D, H, W = vol.shape[-3:]
    return VolumeRaycaster((W, D, H), out_shape, max_samples=M, tf_resolution=R, layout=layout, fov=fov, nearfar=(near, 100.0))


def _cuda_forward(vol, tf, cams, out_shape, jit, M=2048, sr=1.0, nondiff=False, dtype=torch.float32, image_layout=True,
                  layout=DEFAULT_LAYOUT, fov=30.0, near=0.1):
    vr = _vr(vol, out_shape, tf.shape[-1], M, layout, fov, near)
    assert vr.skip_empty
    dev = "cuda:0"
    bricked = vr.brick(vol.to(dev, dtype).reshape(1, *vol.shape[-3:]).contiguous())
    tf_r4 = tf.to(dev).t().contiguous()[None]
    out, K, Tp = vr.march(bricked, tf_r4, cams.to(dev).contiguous(), sr, None if jit is None else jit.to(dev).contiguous(),
                          nondiff=nondiff, image_layout=image_layout)
    return vr, bricked, tf_r4, out, K, Tp


CASES = [
    # (vol_shape(D,H,W), out(w,h), R, views, jitter, sr, M)
    ((32, 32, 32), (48, 40), 16, 2, True, 1.0, 2048),
    ((64, 64, 64), (96, 64), 128, 2, True, 1.0, 2048),
    ((37, 29, 45), (50, 34), 33, 1, False, 1.0, 2048),        # ragged: nothing is a multiple of 4/8
    ((48, 48, 48), (64, 64), 128, 1, True, 0.7, 2048),
    ((40, 40, 40), (40, 40), 64, 1, True, 2.0, 4096),
    ((64, 64, 64), (64, 64), 128, 1, True, 1.0, 40),          # max_samples truncates the rays (H2)
    ((1100, 6, 6), (24, 20), 32, 1, True, 1.0, 4096),         # an axis > 1000 voxels: both +-1e-3 taps of that axis can leave the cell
    ((6, 1100, 6), (24, 20), 32, 1, True, 1.0, 4096),
    ((6, 6, 1100), (24, 20), 32, 1, True, 1.0, 4096),
]


_ORACLE_CACHE = {}


def _oracle_case(case, dtype=torch.float32, tf_name="rand"):
    """Oracle image, counts and gradients of one CASES row (computed once, shared by the layout / dtype variants)."""
    key = (case, dtype, tf_name)
    if key not in _ORACLE_CACHE:
        shape, out_shape, R, views, jitter, sr, M = case
        vol, tf, cams, jit = case_inputs(shape, out_shape, R, seed=len(shape) + R, tf_name=tf_name, views=views, jitter=jitter)
        if dtype == torch.float16:
            vol = vol.half().float()                          # the oracle marches the fp16-rounded values in fp32
        ref, Kr, nr = oracle_forward_views(vol, tf, cams, out_shape, jit, sampling_rate=sr, max_samples=M)
        go = torch.randn(ref.shape, generator=torch.Generator().manual_seed(7))
        gv_ref, gt_ref = oracle_backward_views(vol, tf, cams, go.numpy(), out_shape, jit, sampling_rate=sr, max_samples=M)
        _ORACLE_CACHE[key] = (vol, tf, cams, jit, ref, Kr, go, gv_ref, gt_ref)
    return _ORACLE_CACHE[key]


@pytest.mark.parametrize("layout", ["auto", "linear"])
@pytest.mark.parametrize("case", CASES)
def test_forward_backward_match_oracle(case, layout):
    shape, out_shape, R, views, jitter, sr, M = case
    vol, tf, cams, jit, ref, Kr, go, gv_ref, gt_ref = _oracle_case(case)
    vr, bricked, tf_r4, out, K, Tp = _cuda_forward(vol, tf, cams, out_shape, jit, M=M, sr=sr, layout=layout)
    assert bricked.ndim == (4 if layout == "linear" else 3)               # auto = the cell-major copy at these sizes
    got = out.cpu().numpy()
    same = K.cpu().numpy() == Kr
    # H6: rays whose discrete sample count differs are reported and excluded; they must be (almost) absent
    assert (~same).mean() <= 1e-4, f"{(~same).sum()} rays differ in active sample count"
    diff = np.abs(got - ref)
    assert np.moveaxis(diff, 1, 0)[:, same].max() <= RGBA_TOL          # diff is (views,4,H,W), same is (views,H,W)
    gvol, gtf = vr.march_backward(bricked, tf_r4, cams.cuda().contiguous(), sr, None if jit is None else jit.cuda().contiguous(),
                                  go.cuda().contiguous(), out, K, Tp, True, True)
    assert rel_l2(gvol[0].cpu().numpy(), gv_ref) <= GRAD_TOL
    assert rel_l2(gtf[0].cpu().numpy().T, gt_ref) <= GRAD_TOL


@pytest.mark.parametrize("case", [c for c in CASES if max(c[0]) > 1000])
@pytest.mark.parametrize("tf_name", ["rand", "tf1"])
def test_fp16_cell8_two_neighbour_taps_match_oracle(case, tf_name):
    # the kernel variant behind the C5 numbers: bwd_kernel<__half, cell8, TAPS_TWO, vol, tf, SR1> and the fp16 SKIP forward,
    # against the oracle (not only against the linear layout), both gradients; tf1 has exactly-transparent bins (skip grid, transparent-sample paths)
    shape, out_shape, R, views, jitter, sr, M = case
    vol, tf, cams, jit, ref, Kr, go, gv_ref, gt_ref = _oracle_case(case, torch.float16, tf_name)
    vr, bricked, tf_r4, out, K, Tp = _cuda_forward(vol, tf, cams, out_shape, jit, M=M, sr=sr, dtype=torch.float16, layout="cell8")
    assert bricked.dtype == torch.float16 and bricked.ndim == 3
    same = K.cpu().numpy() == Kr
    assert (~same).mean() <= 1e-4
    assert np.moveaxis(np.abs(out.cpu().numpy() - ref), 1, 0)[:, same].max() <= RGBA_TOL
    gvol, gtf = vr.march_backward(bricked, tf_r4, cams.cuda().contiguous(), sr, jit.cuda().contiguous(), go.cuda().contiguous(),
                                  out, K, Tp, True, True)
    assert rel_l2(gvol[0].cpu().numpy(), gv_ref) <= GRAD_TOL
    assert rel_l2(gtf[0].cpu().numpy().T, gt_ref) <= GRAD_TOL
    # and bit-identical forward without the skip grid
    out2, K2, Tp2 = vr.march(bricked, tf_r4, cams.cuda().contiguous(), sr, jit.cuda().contiguous(), skip=False)
    assert torch.equal(out, out2) and torch.equal(K, K2) and torch.equal(Tp, Tp2)


@pytest.mark.parametrize("tf_layout_4r", [False, True])
def test_top_transfer_function_bin(tf_layout_4r):
    """Intensities at and above 1: the centre bin clamps to R - 1 and BOTH halves of the gradient pair belong to bin
    R - 1 (the reference clamps the upper index, :216-218).  The march writes the pair (lo, lo + 1) into copies padded by one bin and
    tf_reduce_kernel folds the pad into bin R - 1: the last bins of the TF gradient must match the oracle entry by entry."""
    from differender_b200._lib import F_TF_4R
    vol, tf, cams, jit = case_inputs((32, 32, 32), (48, 40), 16, seed=21, views=2)
    vol = (vol * 2.5).clamp(0.0, 1.3)                                          # a good share of the voxels at 1.0 .. 1.3
    assert float((vol >= 1.0).float().mean()) > 0.05
    ref, Kr, _ = oracle_forward_views(vol, tf, cams, (48, 40), jit, max_samples=2048)
    go = torch.randn(ref.shape, generator=torch.Generator().manual_seed(2))
    gv_ref, gt_ref = oracle_backward_views(vol, tf, cams, go.numpy(), (48, 40), jit, max_samples=2048)
    vr, bricked, tf_r4, out, K, Tp = _cuda_forward(vol, tf, cams, (48, 40), jit)
    same = K.cpu().numpy() == Kr
    assert (~same).mean() <= 1e-4 and np.moveaxis(np.abs(out.cpu().numpy() - ref), 1, 0)[:, same].max() <= RGBA_TOL
    c, j = cams.cuda().contiguous(), jit.cuda().contiguous()
    if tf_layout_4r:
        gv, gt = vr.march_backward(bricked, tf.cuda().contiguous()[None], c, 1.0, j, go.cuda().contiguous(), out, K, Tp, True, True, extra_flags=F_TF_4R)
        gt = gt[0].cpu().numpy()
    else:
        gv, gt = vr.march_backward(bricked, tf_r4, c, 1.0, j, go.cuda().contiguous(), out, K, Tp, True, True)
        gt = gt[0].cpu().numpy().T
    assert np.abs(gt_ref[:, -1]).min() > 0                                    # the top bin really receives gradient
    assert np.allclose(gt[:, -3:], gt_ref[:, -3:], rtol=1e-3, atol=1e-6 * np.abs(gt_ref).max())
    assert rel_l2(gt, gt_ref) <= GRAD_TOL and rel_l2(gv[0].cpu().numpy(), gv_ref) <= GRAD_TOL


def test_tf_only_and_volume_only_backward():
    vol, tf, cams, jit = case_inputs((48, 48, 48), (64, 48), 128, seed=3, tf_name="tf1", views=2)
    ref, _, _ = oracle_forward_views(vol, tf, cams, (64, 48), jit, max_samples=2048)
    go = torch.randn(ref.shape, generator=torch.Generator().manual_seed(1))
    gv_ref, gt_ref = oracle_backward_views(vol, tf, cams, go.numpy(), (64, 48), jit, max_samples=2048)
    vr, bricked, tf_r4, out, K, Tp = _cuda_forward(vol, tf, cams, (64, 48), jit)
    args = (bricked, tf_r4, cams.cuda().contiguous(), 1.0, jit.cuda().contiguous(), go.cuda().contiguous(), out, K, Tp)
    gv, gt = vr.march_backward(*args, False, True)
    assert gv is None and rel_l2(gt[0].cpu().numpy().T, gt_ref) <= GRAD_TOL
    gv, gt = vr.march_backward(*args, True, False)
    assert gt is None and rel_l2(gv[0].cpu().numpy(), gv_ref) <= GRAD_TOL


def test_nondiff_matches_oracle():
    vol, tf, cams, _ = case_inputs((64, 64, 64), (80, 72), 128, seed=5, tf_name="tf1", views=1, jitter=False)
    ref, Kr, _ = oracle_forward_views(vol, tf, cams, (80, 72), None, sampling_rate=4.0, nondiff=True)
    _, _, _, out, K, _ = _cuda_forward(vol, tf, cams, (80, 72), None, sr=4.0, nondiff=True)
    same = K.cpu().numpy() == Kr
    assert (~same).mean() <= 1e-4
    assert np.moveaxis(np.abs(out.cpu().numpy() - ref), 1, 0)[:, same].max() <= RGBA_TOL
    assert out.max().item() <= 1.0


def test_fp16_volume_matches_oracle_on_rounded_values():
    vol, tf, cams, jit = case_inputs((48, 48, 48), (64, 48), 128, seed=9, tf_name="tf1", views=1)
    vol16 = vol.half()
    ref, Kr, _ = oracle_forward_views(vol16.float(), tf, cams, (64, 48), jit, max_samples=2048)
    vr, bricked, tf_r4, out, K, Tp = _cuda_forward(vol16, tf, cams, (64, 48), jit, dtype=torch.float16)
    assert bricked.dtype == torch.float16 and bricked.data_ptr() != 0
    same = K.cpu().numpy() == Kr
    assert (~same).mean() <= 1e-4
    assert np.moveaxis(np.abs(out.cpu().numpy() - ref), 1, 0)[:, same].max() <= RGBA_TOL
    go = torch.randn(ref.shape, generator=torch.Generator().manual_seed(2))
    gv_ref, gt_ref = oracle_backward_views(vol16.float(), tf, cams, go.numpy(), (64, 48), jit, max_samples=2048)
    gv, gt = vr.march_backward(bricked, tf_r4, cams.cuda().contiguous(), 1.0, jit.cuda().contiguous(), go.cuda().contiguous(),
                               out, K, Tp, True, True)
    assert gv.dtype == torch.float32
    assert rel_l2(gv[0].cpu().numpy(), gv_ref) <= GRAD_TOL and rel_l2(gt[0].cpu().numpy().T, gt_ref) <= GRAD_TOL


def test_generic_tap_path_equals_corner_reuse_path():
    from differender_b200 import _lib
    vol, tf, cams, jit = case_inputs((40, 40, 40), (48, 48), 64, seed=11, views=1)
    vr, bricked, tf_r4, out, K, Tp = _cuda_forward(vol, tf, cams, (48, 48), jit, layout="linear")    # the generic path reads the linear tensor
    orig = vr.desc

    def desc_generic(*a, **k):
        d = orig(*a, **k)
        d.tap_generic = 1
        return d
    vr.desc = desc_generic
    out2, K2, Tp2 = vr.march(bricked, tf_r4, cams.cuda().contiguous(), 1.0, jit.cuda().contiguous())
    assert torch.equal(K, K2)
    assert torch.equal(out[:, 3], out2[:, 3])                 # the alpha path is bit-identical by construction
    assert (out - out2).abs().max().item() <= 1e-6
    go = torch.randn(out.shape, generator=torch.Generator().manual_seed(4)).cuda()
    a = (bricked, tf_r4, cams.cuda().contiguous(), 1.0, jit.cuda().contiguous(), go, out, K, Tp, True, True)
    gv2, gt2 = vr.march_backward(*a)
    vr.desc = orig
    gv, gt = vr.march_backward(*a)
    assert rel_l2(gv2.cpu().numpy(), gv.cpu().numpy()) <= 1e-4 and rel_l2(gt2.cpu().numpy(), gt.cpu().numpy()) <= 1e-4


@pytest.mark.parametrize("layout,dtype", [("brick8", torch.float32), ("cell8", torch.float32), ("cell8", torch.float16), ("brick8", torch.float16)])
def test_copied_layouts_are_bit_identical_to_linear_layout(layout, dtype):
    vol, tf, cams, jit = case_inputs((37, 29, 45), (56, 40), 64, seed=13, views=2)
    vr, vlin, tf_r4, out, K, Tp = _cuda_forward(vol, tf, cams, (56, 40), jit, dtype=dtype, layout="linear")
    vb, bricked, _, out_b, K_b, Tp_b = _cuda_forward(vol, tf, cams, (56, 40), jit, layout=layout, dtype=dtype)
    # zero-copy view vs bricked / cell-major copy
    assert vlin.shape == (1, 37, 29, 45) and bricked.dtype == dtype
    assert bricked.shape == ((1, 5 * 4 * 6 * 512) if layout == "brick8" else (1, 37 * 29 * 45, 8))
    assert torch.equal(out, out_b) and torch.equal(K, K_b) and torch.equal(Tp, Tp_b)     # same arithmetic, same order
    go = torch.randn(out.shape, generator=torch.Generator().manual_seed(4)).cuda()
    c, j = cams.cuda().contiguous(), jit.cuda().contiguous()
    gv, gt = vr.march_backward(vlin, tf_r4, c, 1.0, j, go, out, K, Tp, True, True)
    gv_b, gt_b = vb.march_backward(bricked, tf_r4, c, 1.0, j, go, out, K, Tp, True, True)
    assert rel_l2(gv_b.cpu().numpy(), gv.cpu().numpy()) <= 1e-5 and rel_l2(gt_b.cpu().numpy(), gt.cpu().numpy()) <= 1e-5
    if dtype == torch.float32:
        ref, Kr, _ = oracle_forward_views(vol, tf, cams, (56, 40), jit, max_samples=2048)
        assert np.array_equal(K_b.cpu().numpy(), Kr) and np.abs(out_b.cpu().numpy() - ref).max() <= RGBA_TOL


@pytest.mark.parametrize("shape", [(1100, 6, 6), (6, 1100, 6), (6, 6, 1100)])
@pytest.mark.parametrize("layout", ["brick8", "cell8"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
def test_copied_layouts_with_both_taps_crossing(shape, layout, dtype):
    vol, tf, cams, jit = case_inputs(shape, (24, 20), 32, seed=5, views=1)
    vr, vlin, tf_r4, out, K, Tp = _cuda_forward(vol, tf, cams, (24, 20), jit, M=4096, dtype=dtype, layout="linear")
    vb, bricked, _, out_b, K_b, Tp_b = _cuda_forward(vol, tf, cams, (24, 20), jit, M=4096, layout=layout, dtype=dtype)
    # the exact path (alpha, K, Tprev) is bit-identical in every layout.  In this two-neighbour regime the cell-major layout
    # evaluates each normal tap from the tap's own cell record (eval_normals_direct): the same mixes on the same values, but a
    # differently shaped kernel in which the compiler fuses the (not exactly rounded) shading arithmetic differently -- RGB may
    # differ in the last bit (measured: <= 6e-8 in a few dozen pixels)
    assert torch.equal(out[:, 3], out_b[:, 3]) and torch.equal(K, K_b) and torch.equal(Tp, Tp_b)
    assert (out - out_b).abs().max().item() <= (2e-7 if layout == "cell8" else 0.0)
    go = torch.randn(out.shape, generator=torch.Generator().manual_seed(4)).cuda()
    c, j = cams.cuda().contiguous(), jit.cuda().contiguous()
    gv, gt = vr.march_backward(vlin, tf_r4, c, 1.0, j, go, out, K, Tp, True, True)
    gv_b, gt_b = vb.march_backward(bricked, tf_r4, c, 1.0, j, go, out, K, Tp, True, True)
    assert rel_l2(gv_b.cpu().numpy(), gv.cpu().numpy()) <= 1e-5 and rel_l2(gt_b.cpu().numpy(), gt.cpu().numpy()) <= 1e-5


def test_raw_layout_equals_image_layout():
    vol, tf, cams, jit = case_inputs((32, 32, 32), (40, 24), 16, seed=2, views=2)
    _, _, _, img, K, _ = _cuda_forward(vol, tf, cams, (40, 24), jit)
    _, _, _, raw, K2, _ = _cuda_forward(vol, tf, cams, (40, 24), jit, image_layout=False)
    # reference L3: torch.flip(raw, (2,)).permute(0, 3, 2, 1)   (volume_raycaster.py:538-541)
    assert torch.equal(torch.flip(raw, (2,)).permute(0, 3, 2, 1).contiguous(), img)
    assert torch.equal(K, K2)


def test_golden_fixtures_on_gpu():
    import glob, os
    files = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))
    assert files, "no golden fixtures committed"
    for f in files:
        z = np.load(f)
        vol, tf, cams = torch.tensor(z["volume"]), torch.tensor(z["tf"]), torch.tensor(z["cams"])
        jit = torch.tensor(z["jitter"]) if "jitter" in z.files else None
        out_shape = tuple(int(v) for v in z["output_shape"])
        sr, M = float(z["sampling_rate"]), int(z["max_samples"])
        vr, bricked, tf_r4, out, K, Tp = _cuda_forward(vol, tf, cams, out_shape, jit, M=M, sr=sr)
        same = K.cpu().numpy() == z["K"]
        assert (~same).mean() <= 1e-4
        d = np.abs(out.cpu().numpy() - z["image"])
        assert np.moveaxis(d, 1, 0)[:, same].max() <= RGBA_TOL, f
        go = torch.tensor(z["grad_image"])
        gv, gt = vr.march_backward(bricked, tf_r4, cams.cuda().contiguous(), sr, None if jit is None else jit.cuda().contiguous(),
                                   go.cuda().contiguous(), out, K, Tp, True, True)
        assert rel_l2(gv[0].cpu().numpy(), z["grad_volume"]) <= GRAD_TOL, f
        assert rel_l2(gt[0].cpu().numpy().T, z["grad_tf"]) <= GRAD_TOL, f


def test_tf_in_torch_4xR_layout_equals_Rx4_layout():
    from differender_b200._lib import F_TF_4R
    vol, tf, cams, jit = case_inputs((32, 32, 32), (40, 24), 48, seed=17, tf_name="tf1", views=2)
    vr, v, tf_r4, out, K, Tp = _cuda_forward(vol, tf, cams, (40, 24), jit)
    tf_4r = tf.cuda().contiguous()[None]                                       # [1, 4, R]: the torch-side layout, no transpose copy
    c, j = cams.cuda().contiguous(), jit.cuda().contiguous()
    out2, K2, Tp2 = vr.march(v, tf_4r, c, 1.0, j, extra_flags=F_TF_4R)
    assert torch.equal(out, out2) and torch.equal(K, K2)
    go = torch.randn(out.shape, generator=torch.Generator().manual_seed(3)).cuda()
    gv, gt = vr.march_backward(v, tf_r4, c, 1.0, j, go, out, K, Tp, True, True)
    gv2, gt2 = vr.march_backward(v, tf_4r, c, 1.0, j, go, out, K, Tp, True, True, extra_flags=F_TF_4R)
    assert gt2.shape == (1, 4, 48)
    assert rel_l2(gt2[0].t().cpu().numpy(), gt[0].cpu().numpy()) <= 1e-5 and rel_l2(gv2.cpu().numpy(), gv.cpu().numpy()) <= 1e-5
