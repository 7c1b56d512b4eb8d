"""The oracle pinned against the REFERENCE'S OWN SOURCE.

tests/golden/shim/*.npz hold what differender/volume_raycaster.py computes when its source is executed on oracle/ti_shim.py (a
strict-IEEE-fp32 interpreter of the Taichi subset it uses; tests/golden/make_shim_golden.py).  The oracle build without any
contraction (`source_order`) must reproduce them BIT FOR BIT (image, sample counts, early-termination counts) and to float64
accumulation accuracy on both gradients; the default oracle build (the contractions it defines, oracle/cpu_ref.c header) and the CUDA
path must stay inside the north_star tolerances.  What this does not cover -- the real Taichi compiler's rounding -- is what
tools/rounding_envelope.py bounds."""
import glob
import os

import numpy as np
import pytest
import torch

from helpers import GRAD_TOL, RGBA_TOL
from oracle import cpu_oracle as co
from oracle import taichi_probe as tp
from oracle import ti_shim

FIXTURES = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "shim", "*.npz")))


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def _rel(a, b, nan):
    a, b = np.where(nan, 0, np.asarray(a, np.float64)), np.where(nan, 0, np.asarray(b, np.float64))
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


def _kw(z, variant):
    kw = dict(sampling_rate=float(z["sampling_rate"]), max_samples=int(z["max_samples"]), fov=float(z["fov"]), near=float(z["near"]), variant=variant)
    if "jitter" in z.files:
        kw["jitter"] = z["jitter"]
    return kw


def test_fixtures_are_committed():
    assert len(FIXTURES) >= 7
    kinds = [bool(np.load(f)["nondiff"]) for f in FIXTURES]
    assert any(kinds) and not all(kinds)


@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(f)[:-4] for f in FIXTURES])
def test_source_order_oracle_is_bit_identical_to_the_reference_source(path):
    z = np.load(path)
    res = tuple(int(v) for v in z["output_shape"])
    live = z["n"] > 1                                         # n == 1 rays are 0/0 in the reference
    if bool(z["nondiff"]):
        img, _, n = co.forward(z["volume"], z["tf"], z["cam"], res, return_counts=True, nondiff=True, **_kw(z, "source_order"))
        assert np.array_equal(n, z["n"])
        assert np.array_equal(_bits(img)[:, live], _bits(z["image"])[:, live])
        return
    img, K, n = co.forward(z["volume"], z["tf"], z["cam"], res, return_counts=True, **_kw(z, "source_order"))
    assert np.array_equal(n, z["n"]) and np.array_equal(K[live], z["K"][live])
    assert np.array_equal(_bits(img)[:, live], _bits(z["image"])[:, live])
    gv, gt = co.backward(z["volume"], z["tf"], z["cam"], z["grad_image"], res, **_kw(z, "source_order"))
    # the interpreter accumulates adjoints in float64, the oracle forms each sample's adjoint in fp32
    assert _rel(gv, z["grad_volume"], z["gvol_nan"]) <= 5e-6 and _rel(gt, z["grad_tf"], z["gtf_nan"]) <= 5e-7
    # H4: where the reference poisons a voxel's gradient with NaN (and then zeroes it), the oracle keeps the finite contributions
    assert np.isfinite(gv).all() and np.isfinite(gt).all()
    assert (z["grad_volume"][z["gvol_nan"]] == 0).all()


@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(f)[:-4] for f in FIXTURES])
def test_default_oracle_is_within_tolerance_of_the_reference_source(path):
    z = np.load(path)
    res = tuple(int(v) for v in z["output_shape"])
    live = z["n"] > 1
    nd = bool(z["nondiff"])
    img, K, n = co.forward(z["volume"], z["tf"], z["cam"], res, return_counts=True, nondiff=nd, **_kw(z, None))
    assert np.array_equal(n, z["n"])
    assert np.abs(img - z["image"])[:, live].max() <= RGBA_TOL
    if nd:
        return
    assert np.array_equal(K[live], z["K"][live])
    gv, gt = co.backward(z["volume"], z["tf"], z["cam"], z["grad_image"], res, **_kw(z, None))
    assert _rel(gv, z["grad_volume"], z["gvol_nan"]) <= GRAD_TOL and _rel(gt, z["grad_tf"], z["gtf_nan"]) <= GRAD_TOL


def test_shim_regenerates_a_fixture_from_the_mounted_reference():
    """Only where the reference tree is mounted (the development container): the committed fixture IS what the reference source
    computes on the interpreter."""
    if tp.find_reference() is None:
        pytest.skip("reference source not mounted (GPU box)")
    z = np.load(FIXTURES[0])
    c = tp.SHIM_CASES[0]
    assert str(z["name"]) == c["name"]
    mod = tp._load_reference_module("shim")
    vol, tf, cam, jit, go = tp.case_inputs(c)
    assert np.array_equal(vol, z["volume"]) and np.array_equal(tf, z["tf"]) and np.array_equal(jit, z["jitter"])
    ti_shim.reset()
    r = tp.run_reference(mod, vol, tf, cam, jit, go, c["res"], c["M"], c["sr"])
    assert np.array_equal(_bits(r["image"]), _bits(z["image"])) and np.array_equal(r["K"], z["K"]) and np.array_equal(r["n"], z["n"])
    assert np.array_equal(r["gvol"], z["grad_volume"]) and np.array_equal(r["gtf"], z["grad_tf"])
    assert np.array_equal(r["gvol_nan"], z["gvol_nan"])


# ------------------------------------------------------------------------------------------------ the interpreter itself
def test_shim_scalars_round_once_to_fp32_per_operator():
    F, f32 = ti_shim.F, np.float32
    a, b, c = 0.1, 0.7, 1e-3
    x = F(f32(a)) * b + c                                     # Python constants become fp32 when they meet an fp32 value
    assert x.v == f32(f32(f32(a) * f32(b)) + f32(c)) and x.v.dtype == np.float32
    assert (F(f32(1)) / 3).v == f32(1) / f32(3)
    assert ti_shim.tan(0.5) == np.tan(0.5) and isinstance(ti_shim.tan(0.5), float)      # Python scope: stays a double
    assert ti_shim.ti_int(F(f32(-2.7))) == -2 and ti_shim.ti_int(F(f32(2.7))) == 2    # casts truncate
    with np.errstate(all="ignore"):                           # outside a kernel call the interpreter leaves numpy's error state alone
        nan = F(f32(np.nan))
        assert ti_shim.ti_max(nan, 0.0).v == 0 and np.isnan(ti_shim.ti_max(0.0, nan).v)     # a > b ? a : b
        assert ti_shim.ti_min(nan, 1.0).v == 1 and np.isnan(ti_shim.ti_min(1.0, nan).v)
        v = ti_shim.Vec([F(f32(3)), F(f32(4)), F(f32(0))])
        assert v.norm().v == 5 and [e.v for e in v.normalized().e] == [f32(f32(1) / f32(5)) * f32(3), f32(f32(1) / f32(5)) * f32(4), 0]
        assert np.isnan(ti_shim.Vec([F(f32(0))] * 3).normalized().x.v)                     # 0 * inf


def _kernel_fixture():
    ti, tl = ti_shim.make_modules()
    ti_shim.reset()

    class K:
        def __init__(self):
            self.x = ti.field(ti.f32, shape=4, needs_grad=True)
            self.t = ti.Vector.field(4, dtype=ti.f32, shape=2, needs_grad=True)
            self.y = ti.field(ti.f32, shape=(), needs_grad=True)

        @ti.kernel
        def f(self, p: ti_shim.ti_float):
            a = tl.mix(self.x[0], self.x[1], 0.3) * ti.sqrt(self.x[2]) / self.x[3]
            c = tl.mix(self.t[0], self.t[1], a)
            n = tl.vec3(c.x, c.y, c.z).normalized()
            self.y[None] = ti.pow(ti.max(n.dot(tl.vec3(0.2, -0.5, 0.7)), 0.0), p) + ti.min(1.0, c.w * a) + ti.floor(a)
    return K()


def test_shim_reverse_sweep_matches_finite_differences():
    k = _kernel_fixture()
    x0 = np.array([0.3, 0.9, 0.8, 1.7], np.float32)
    t0 = np.array([[0.2, 0.5, 0.1, 0.4], [0.9, 0.3, 0.6, 0.8]], np.float32)

    def run(x, t):
        k.x.from_torch(torch.tensor(x)); k.t.from_torch(torch.tensor(t))
        k.f(1.7)
        return float(k.y.to_torch())
    run(x0, t0)
    k.x.grad.fill(0); k.t.grad.fill(0); k.y.grad.fill(0)
    k.y.grad.from_torch(torch.tensor(1.0))
    k.f.grad(1.7)
    gx, gt = k.x.grad.to_numpy64(), k.t.grad.to_numpy64()
    assert np.abs(gx).min() > 0 and np.abs(gt).min() > 0
    h = 2e-3                                                  # fp32 forward: central differences are good to ~1e-3 relative
    for i in range(4):
        d = np.zeros(4, np.float32); d[i] = h
        fd = (run(x0 + d, t0) - run(x0 - d, t0)) / (float((x0 + d)[i]) - float((x0 - d)[i]))
        assert abs(fd - gx[i]) <= 5e-3 * max(1.0, abs(gx[i])), (i, fd, gx[i])
    for i in range(2):
        for c in range(4):
            d = np.zeros((2, 4), np.float32); d[i, c] = h
            fd = (run(x0, t0 + d) - run(x0, t0 - d)) / (float((t0 + d)[i, c]) - float((t0 - d)[i, c]))
            assert abs(fd - gt[i, c]) <= 5e-3 * max(1.0, abs(gt[i, c])), (i, c, fd, gt[i, c])


def test_shim_adjoint_conventions_of_the_reference_autodiff():
    """max/min route the whole adjoint to the selected operand (a tie goes to the right-hand one); a zero adjoint still multiplies an
    infinite partial (the NaN poisoning of SURVEY 7.3 H4)."""
    ti, tl = ti_shim.make_modules()
    ti_shim.reset()

    class K:
        def __init__(self):
            self.x = ti.field(ti.f32, shape=3, needs_grad=True)
            self.y = ti.field(ti.f32, shape=2, needs_grad=True)

        @ti.kernel
        def f(self):
            self.y[0] = ti.max(self.x[0], 0.0) + ti.max(self.x[1], self.x[1] * 1.0)
            v = tl.vec3(self.x[2], self.x[2], self.x[2]).normalized()          # x[2] = 0: 0/0
            self.y[1] = ti.max(v.x, 0.0) + self.x[0]
    k = K()
    k.x.from_torch(torch.tensor([0.0, 2.0, 0.0]))
    k.f()
    assert k.y.to_torch().tolist() == [2.0, 0.0]              # max(NaN, 0) = 0 launders the forward
    k.y.grad.from_torch(torch.tensor([1.0, 1.0]))
    k.f.grad()
    g = k.x.grad.to_numpy64()
    assert g[0] == 1.0                                        # tie max(0, 0.0): the constant takes it; + the direct use in y[1]
    assert g[1] == 1.0                                        # tie max(a, a*1): the right-hand operand, which is a*1
    assert np.isnan(g[2])                                     # adjoint 0 x partial inf


# ------------------------------------------------------------------------------------------------------- the CUDA path
@pytest.mark.gpu
@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(f)[:-4] for f in FIXTURES])
def test_cuda_path_against_the_reference_source(path):
    from test_gpu_parity import _cuda_forward
    z = np.load(path)
    res = tuple(int(v) for v in z["output_shape"])
    sr, M, nd = float(z["sampling_rate"]), int(z["max_samples"]), bool(z["nondiff"])
    vol, tf, cams = torch.tensor(z["volume"])[None], torch.tensor(z["tf"]), torch.tensor(z["cam"])[None]
    jit = torch.tensor(z["jitter"])[None] if "jitter" in z.files else None
    vr, bricked, tf_r4, out, K, Tp = _cuda_forward(vol, tf, cams, res, jit, M=M, sr=sr, nondiff=nd, fov=float(z["fov"]), near=float(z["near"]))
    live = z["n"] > 1
    assert np.abs(out[0].cpu().numpy() - z["image"])[:, live].max() <= RGBA_TOL
    if nd:
        return
    assert np.array_equal(K[0].cpu().numpy()[live], z["K"][live])
    go = torch.tensor(z["grad_image"])[None].cuda().contiguous()
    gv, gt = vr.march_backward(bricked, tf_r4, cams.cuda().contiguous(), sr, None if jit is None else jit.cuda().contiguous(),
                               go, out, K, Tp, True, True)
    assert _rel(gv[0].cpu().numpy(), z["grad_volume"], z["gvol_nan"]) <= GRAD_TOL
    assert _rel(gt[0].cpu().numpy().T, z["grad_tf"], z["gtf_nan"]) <= GRAD_TOL


# ------------------------------------------------------------------------------- the reference's PUBLIC API on the interpreter
API_FIXTURES = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "shim_api", "*.npz")))


def _api_items(z):
    """Per batch item: (volume (D,H,W), tf (4,R), cam (3,), image (4,H,W), grad_output (4,H,W), nondiff image)."""
    vol, tf, look, img, go, nd = z["volume"], z["tf"], z["look_from"], z["image"], z["grad_output"], z["image_nondiff"]
    bs = img.shape[0] if img.ndim == 4 else 1
    for b in range(bs):
        yield (vol[b, 0] if vol.ndim == 5 else vol[0], tf[b] if tf.ndim == 3 else tf, look[b] if look.ndim == 2 else look,
               img[b] if img.ndim == 4 else img, go[b] if go.ndim == 4 else go, nd[b] if nd.ndim == 4 else nd)


def test_api_fixtures_are_committed():
    assert len(API_FIXTURES) >= 3
    nd = [np.load(f)["image"].ndim for f in API_FIXTURES]
    assert 3 in nd and 4 in nd                                # non-batched and batched calls


@pytest.mark.parametrize("path", API_FIXTURES, ids=[os.path.basename(f)[:-4] for f in API_FIXTURES])
def test_oracle_matches_the_reference_public_api_on_the_interpreter(path):
    """`Raycaster(...)(volume, tf, look_from)` + autograd and `raycast_nondiff` of the reference (its _determine_batch, flips and
    permutes included) against the un-contracted oracle build called per item with the user-level tensors."""
    z = np.load(path)
    res = tuple(int(v) for v in z["output_shape"])
    sr, M = float(z["sampling_rate"]), int(z["max_samples"])
    J = z["jitter"] if "jitter" in z.files else None
    gvs, gts = [], []
    for vol, tf, cam, img, go, nd in _api_items(z):
        o, K, n = co.forward(vol, tf, cam, res, return_counts=True, sampling_rate=sr, max_samples=M, jitter=J, variant="source_order")
        live = n > 1
        assert np.array_equal(_bits(o)[:, live], _bits(img)[:, live])
        o_nd, _, n_nd = co.forward(vol, tf, cam, res, return_counts=True, sampling_rate=4.0 * sr, max_samples=M, nondiff=True, variant="source_order")
        assert np.array_equal(_bits(o_nd)[:, n_nd > 1], _bits(nd)[:, n_nd > 1])                 # :503: 4 x the sampling rate, no jitter
        gv, gt = co.backward(vol, tf, cam, go, res, sampling_rate=sr, max_samples=M, jitter=J, variant="source_order")
        gvs.append(gv); gts.append(gt)
    none = np.zeros((), bool)
    gV = np.stack(gvs)[:, None] if z["volume"].ndim == 5 else sum(gvs)[None]                    # shared inputs: autograd sums over the batch
    gT = np.stack(gts) if z["tf"].ndim == 3 else sum(gts)
    assert gV.shape == z["grad_volume"].shape and gT.shape == z["grad_tf"].shape
    # The reference NaN-poisons and then zeroes the voxels an n == 1 ray (0/0 position -> voxel (0,0,0) and its neighbours, H3/H4)
    # touches, even with a zero seed; the oracle keeps their finite gradient.  Those few corner voxels are excluded.
    poisoned = (z["grad_volume"] == 0) & (gV != 0)
    assert poisoned.sum() <= 8 * len(gvs)
    assert _rel(gV, z["grad_volume"], poisoned) <= 5e-6 and _rel(gT, z["grad_tf"], none) <= 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("path", API_FIXTURES, ids=[os.path.basename(f)[:-4] for f in API_FIXTURES])
def test_drop_in_raycaster_against_the_reference_public_api(path):
    """The same constructor and the same calls on `differender_b200.Raycaster` (CUDA) and on the reference's `Raycaster` (its source
    on the interpreter): images, nondiff images and the gradients autograd hands back for the user's tensors."""
    from differender_b200 import Raycaster
    z = np.load(path)
    res = tuple(int(v) for v in z["output_shape"])
    shape = tuple(int(v) for v in z["volume_shape"])
    sr, M = float(z["sampling_rate"]), int(z["max_samples"])
    J = torch.tensor(z["jitter"]).cuda() if "jitter" in z.files else None
    rc = Raycaster(shape, res, z["tf"].shape[-1], sampling_rate=sr, jitter=J is not None, max_samples=M)
    vol = torch.tensor(z["volume"]).cuda().requires_grad_(True)
    tf = torch.tensor(z["tf"]).cuda().requires_grad_(True)
    look = torch.tensor(z["look_from"]).cuda()
    img = rc(vol, tf, look, jitter_tensor=J)
    assert tuple(img.shape) == z["image"].shape
    live = np.ones(z["image"].shape, bool)
    items = list(_api_items(z))
    for b, (v, t, cam, _, _, _) in enumerate(items):          # mask the n <= 1 rays (H3)
        _, _, n = co.forward(v, t, cam, res, return_counts=True, sampling_rate=sr, max_samples=M, jitter=None if J is None else z["jitter"])
        (live[b] if z["image"].ndim == 4 else live)[:, n <= 1] = False
    assert np.abs(img.detach().cpu().numpy() - z["image"])[live].max() <= RGBA_TOL
    (img * torch.tensor(z["grad_output"]).cuda()).sum().backward()
    none = np.zeros((), bool)
    assert tuple(vol.grad.shape) == z["grad_volume"].shape and tuple(tf.grad.shape) == z["grad_tf"].shape
    gV = vol.grad.cpu().numpy()
    poisoned = (z["grad_volume"] == 0) & (gV != 0)            # H3/H4: zeroed by the reference's nan_to_num, kept finite here
    assert poisoned.sum() <= 8 * len(items)
    assert _rel(gV, z["grad_volume"], poisoned) <= GRAD_TOL
    assert _rel(tf.grad.cpu().numpy(), z["grad_tf"], none) <= GRAD_TOL
    nd = rc.raycast_nondiff(vol.detach(), tf.detach(), look)
    assert tuple(nd.shape) == z["image_nondiff"].shape
    assert np.abs(nd.cpu().numpy() - z["image_nondiff"])[live].max() <= RGBA_TOL


# ----------------------------------------------------- the device arithmetic (csrc/dr_math.cuh compiled for the host) on the CPU
@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(f)[:-4] for f in FIXTURES])
@pytest.mark.parametrize("layout", ["linear", "cell8"])
def test_device_math_against_the_reference_source(path, layout):
    """The kernels' per-ray code, run on the host by tests/hostsim, against what the reference source computes: the same bar the
    GPU test applies, checkable where there is no GPU."""
    import hostsim_lib as hs
    z = np.load(path)
    res = tuple(int(v) for v in z["output_shape"])
    kw = dict(sampling_rate=float(z["sampling_rate"]), max_samples=int(z["max_samples"]), fov=float(z["fov"]), near=float(z["near"]),
              jitter=z["jitter"] if "jitter" in z.files else None, cell=layout == "cell8")
    live = z["n"] > 1
    out, K, _, n = hs.forward(z["volume"], z["tf"], z["cam"], res, nondiff=bool(z["nondiff"]), **kw)
    assert np.array_equal(n, z["n"])
    assert np.abs(out - z["image"])[:, live].max() <= RGBA_TOL
    if bool(z["nondiff"]):
        return
    assert np.array_equal(K[live], z["K"][live])
    gv, gt = hs.backward(z["volume"], z["tf"], z["cam"], z["grad_image"], res, **kw)
    assert _rel(gv, z["grad_volume"], z["gvol_nan"]) <= GRAD_TOL and _rel(gt, z["grad_tf"], z["gtf_nan"]) <= GRAD_TOL


# ------------------------------------------------- the reference's TF-optimisation DEMO loop (BASELINE config C2) on the interpreter
LOOP_FIXTURE = os.path.join(os.path.dirname(__file__), "golden", "shim_loop", "l0_tf_loop.npz")


def test_oracle_reproduces_the_reference_demo_loop():
    """Reference code only, on the interpreter (make_shim_golden.py main_loop): the library's kernels driven in the order of the demo's
    `backward()` (examples/taichi_volume_raycaster.py:425-447: jittered march at sr 0.7, F.mse_loss, TF gradient) and the demo's own
    `apply_grad` kernel (:375-381) with its learning-rate decay, three iterations from `black` towards a `tf1` render.  The same loop with
    the oracle (un-contracted build) and oracle/aux_ref.py."""
    from oracle import aux_ref
    z = np.load(LOOP_FIXTURE)
    res = tuple(int(v) for v in z["output_shape"])
    M, sr = int(z["max_samples"]), float(z["bw_sampling_rate"])
    tgt = co.forward(z["volume"], z["tf_target"], z["cam"], res, sampling_rate=float(z["fw_sampling_rate"]), max_samples=M, nondiff=True,
                     variant="source_order")
    assert np.array_equal(_bits(tgt), _bits(z["target"]))                      # the target image (the library's nondiff march), bit for bit
    tf, mom, lr = z["tf_init"].copy(), np.zeros_like(z["tf_init"]), float(z["lr"])
    for k in range(z["tf_after"].shape[0]):
        kw = dict(sampling_rate=sr, max_samples=M, jitter=z["jitter"][k], variant="source_order")
        out, _, n = co.forward(z["volume"], tf, z["cam"], res, return_counts=True, **kw)
        assert (n != 1).all()                                                 # no 0/0 rays in this case (H3)
        d = out.astype(np.float32) - z["target"]
        loss = float((d.astype(np.float64) ** 2).mean())
        assert abs(loss - float(z["loss"][k])) <= 1e-6 * float(z["loss"][k])
        go = (d * np.float32(2.0 / d.size)).astype(np.float32)                  # d mse_loss / d out
        _, gt = co.backward(z["volume"], tf, z["cam"], go, res, want_vol=False, **kw)
        assert _rel(gt, z["grad_tf"][k], np.zeros((), bool)) <= 2e-6
        tf, mom = aux_ref.momentum_step(tf, gt.astype(np.float32), mom, lr, float(z["momentum"]), float(z["clip"]))
        lr *= float(z["lr_decay"])
        assert np.abs(tf - z["tf_after"][k]).max() <= 1e-7, k                   # the transfer function after each demo iteration
    assert np.abs(z["tf_after"][-1] - z["tf_init"]).max() > 1e-2                # and the loop really moved it


@pytest.mark.gpu
def test_drop_in_loop_against_the_reference_demo_loop():
    """The same three iterations through the product: Raycaster.mse_loss (render + MSE fused) + backward + MomentumSGD.step."""
    from differender_b200 import MomentumSGD, Raycaster
    z = np.load(LOOP_FIXTURE)
    res = tuple(int(v) for v in z["output_shape"])
    D, H, W = z["volume"].shape
    rc = Raycaster((D, H, W), res, z["tf_init"].shape[1], sampling_rate=float(z["bw_sampling_rate"]), jitter=True, max_samples=int(z["max_samples"]))
    vol = torch.tensor(z["volume"])[None].cuda()
    cam = torch.tensor(z["cam"]).cuda()
    nd = rc.raycast_nondiff(vol, torch.tensor(z["tf_target"]).cuda(), cam, sampling_rate=float(z["fw_sampling_rate"]))
    assert np.abs(nd.cpu().numpy() - z["target"]).max() <= RGBA_TOL
    target = torch.tensor(z["target"]).cuda()
    tf = torch.tensor(z["tf_init"]).cuda().requires_grad_(True)
    opt = MomentumSGD(tf, lr=float(z["lr"]), momentum=float(z["momentum"]), max_grad=float(z["clip"]), lr_decay=float(z["lr_decay"]))
    for k in range(z["tf_after"].shape[0]):
        tf.grad = None
        loss, _ = rc.mse_loss(vol, tf, cam, target, jitter_tensor=torch.tensor(z["jitter"][k]).cuda())
        loss.backward()
        assert abs(float(loss.detach()) - float(z["loss"][k])) <= 1e-4 * float(z["loss"][k])
        assert _rel(tf.grad.cpu().numpy(), z["grad_tf"][k], np.zeros((), bool)) <= GRAD_TOL
        opt.step()
        assert np.abs(tf.detach().cpu().numpy() - z["tf_after"][k]).max() <= 1e-5, k


def test_determine_batch_matches_the_reference_rule():
    """`Raycaster._determine_batch` (:551-571) of the reference source against the product's, for every combination of batched and
    shared inputs: same batched flag, same batch size, same values (the product hands out views of shared inputs where the reference
    expands and clones them)."""
    if tp.find_reference() is None:
        pytest.skip("reference source not mounted (GPU box)")
    from differender_b200 import Raycaster
    mod = tp._load_reference_module("shim")
    ref_self, my_self = object.__new__(mod.Raycaster), object.__new__(Raycaster)
    g = torch.Generator().manual_seed(3)
    BS, D, H, W, R = 3, 4, 5, 6, 7
    for bv in (False, True):
        for bt in (False, True):
            for bl in (False, True):
                vol = torch.rand((BS, 1, D, H, W) if bv else (1, D, H, W), generator=g)
                tf = torch.rand((BS, 4, R) if bt else (4, R), generator=g)
                lf = torch.rand((BS, 3) if bl else (3,), generator=g)
                rb, rbs, rv, rt, rl = mod.Raycaster._determine_batch(ref_self, vol, tf, lf)
                mb, mbs, mv, mt, ml = Raycaster._determine_batch(my_self, vol, tf, lf)
                assert bool(rb) == bool(mb) == (bv or bt or bl) and int(rbs) == int(mbs)
                if rb:                                           # broadcast the product's shared views to the reference's cloned batch
                    mv = mv if mv.ndim == 4 else mv.expand(BS, -1, -1, -1)
                    mt = mt if mt.ndim == 3 else mt.expand(BS, -1, -1)
                    ml = ml if ml.ndim == 2 else ml.expand(BS, -1)
                assert torch.equal(rv, mv.contiguous()) and torch.equal(rt, mt.contiguous()) and torch.equal(rl, ml.contiguous()), (bv, bt, bl)
