#!/usr/bin/env python
"""tests/golden/shim/*.npz: outputs of the REFERENCE'S OWN SOURCE (differender/volume_raycaster.py, read from /root/reference or
$DIFFERENDER_REFERENCE, UNMODIFIED: its ti.random() jitter is served from the supplied jitter tensor) executed on oracle/ti_shim.py, the strict-IEEE-fp32 interpreter of
the Taichi subset it uses.  Run in the development container (the reference tree does not travel to the GPU box):

    python tests/golden/make_shim_golden.py

Each file holds the inputs (volume (D,H,W), tf (4,R), cam (3,), jitter (H,W) if any, grad_image (4,H,W) with the seeds of n <= 1 rays
zeroed), the parameters, and what the reference source computed: image (4,H,W), K, n (H,W), grad_volume (D,H,W) / grad_tf (4,R) in
float64 after its nan_to_num, and the masks of the entries it had NaN-poisoned before that.  Forward-only
(`nondiff`) cases hold image and n.

tests/golden/shim_api/*.npz: the same, one level up -- the reference's PUBLIC API (`Raycaster(...)(volume, tf, look_from)`, its
`_determine_batch`, `RaycastFunction` through torch.autograd, the flips and permutes of :525-548, and `raycast_nondiff`) executed on
the interpreter, non-batched and batched, with the user-level tensors: volume ([BS,] 1, D, H, W), tf ([BS,] 4, R), look_from
([BS,] 3), image ([BS,] 4, H, W) and the gradients autograd returns for them."""
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
        fov, near = c.get("fov", 30.0), c.get("near", 0.1)
        r = tp.run_reference(mod, vol, tf, cam, None if nondiff else jit, None if nondiff else go, c["res"], c["M"], c["sr"], nondiff=nondiff,
                             fov=fov, near=near)
        z = dict(name=c["name"], volume=vol, tf=tf, cam=cam, output_shape=np.array(c["res"]), sampling_rate=c["sr"], max_samples=c["M"],
                 fov=fov, near=near, nondiff=nondiff, image=r["image"], n=r["n"])
        if not nondiff:
            if jit is not None:
                z["jitter"] = jit
            z.update(grad_image=r["grad_image"], K=r["K"], grad_volume=r["gvol"], grad_tf=r["gtf"], gvol_nan=r["gvol_nan"], gtf_nan=r["gtf_nan"])
        slug = re.sub(r"[^a-z0-9]+", "_", c["name"].lower()).strip("_")
        path = os.path.join(out_dir, f"s{k}_{slug}.npz")
        np.savez_compressed(path, **z)
        print(f"{path}: {os.path.getsize(path)} bytes, rays with samples {int((r['n'] > 0).sum())}, longest {int(r['n'].max())}")


API_CASES = [
    dict(name="single 12x16x8 tf1 16x8", shape=(12, 16, 8), tf="tf1", R=32, res=(16, 8), M=64, sr=1.0, jitter=True, cams=[2]),
    dict(name="batched cameras 12^3 rand 8x16 sr 0.7", shape=(12, 12, 12), tf="rand", R=16, res=(8, 16), M=64, sr=0.7, jitter=True, cams=[1, 6, 12]),
    dict(name="batched volumes and tfs 8x12x12 tf3 8x8 no jitter", shape=(8, 12, 12), tf="tf3", R=24, res=(8, 8), M=64, sr=1.0, jitter=False,
         cams=[4, 9], batch_all=True),
]


def main_api():
    import warnings
    import torch
    from differender_b200.synthetic import make_cameras
    from oracle import cpu_oracle as co
    mod = tp._load_reference_module("shim")
    out_dir = os.path.join(HERE, "shim_api")
    os.makedirs(out_dir, exist_ok=True)
    for k, c in enumerate(API_CASES):
        vol, tf, _, jit, _ = tp.case_inputs(c)
        D, H, W = vol.shape
        w, h = c["res"]
        cams = make_cameras(16)[c["cams"]].numpy()
        bs = len(c["cams"])
        ti_shim.reset()
        rc = mod.Raycaster((D, H, W), c["res"], tf.shape[1], sampling_rate=c["sr"], jitter=c["jitter"], max_samples=c["M"])
        if jit is not None:
            tp.set_jitter(rc.vr, jit, copies=2 * bs + 2)                                       # every item marches it, forward and (batched, :456) backward re-run
        if c.get("batch_all"):                                                                 # every input batched, items differ
            volume = torch.tensor(np.stack([vol, np.clip(vol[::-1].copy() * 0.9 + 0.05, 0, 1)]))[:, None]
            tft = torch.tensor(np.stack([tf, tf[:, ::-1].copy()]))
            look = torch.tensor(cams)
        else:
            volume, tft = torch.tensor(vol)[None], torch.tensor(tf)
            look = torch.tensor(cams[0]) if bs == 1 else torch.tensor(cams)
        volume.requires_grad_(True); tft.requires_grad_(True)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            image = rc(volume, tft, look)
        # rays with n <= 1 (0/0 sample position in the reference) get no gradient seed; n from the bit-identical oracle build
        go = np.random.default_rng(11).standard_normal(tuple(image.shape)).astype(np.float32)
        for b in range(bs):
            v = volume[b, 0] if volume.ndim == 5 else volume[0]
            t = tft[b] if tft.ndim == 3 else tft
            _, _, n = co.forward(v.detach().numpy(), t.detach().numpy(), cams[b], c["res"], return_counts=True, sampling_rate=c["sr"],
                                 max_samples=c["M"], jitter=jit, variant="source_order")
            if image.ndim == 4:
                go[b][:, n <= 1] = 0
            else:
                go[:, n <= 1] = 0
        (image * torch.tensor(go)).sum().backward()
        with torch.no_grad():
            nd = rc.raycast_nondiff(volume.detach(), tft.detach(), look)
        z = dict(name=c["name"], volume=volume.detach().numpy(), tf=tft.detach().numpy(), look_from=look.numpy(), output_shape=np.array(c["res"]),
                 volume_shape=np.array((D, H, W)), sampling_rate=c["sr"], max_samples=c["M"], image=image.detach().numpy(), grad_output=go,
                 grad_volume=volume.grad.numpy(), grad_tf=tft.grad.numpy(), image_nondiff=nd.numpy())
        if jit is not None:
            z["jitter"] = jit
        slug = re.sub(r"[^a-z0-9]+", "_", c["name"].lower()).strip("_")
        path = os.path.join(out_dir, f"a{k}_{slug}.npz")
        np.savez_compressed(path, **z)
        print(f"{path}: {os.path.getsize(path)} bytes, image {tuple(image.shape)}, grad_volume {tuple(volume.grad.shape)}, grad_tf {tuple(tft.grad.shape)}")


EXAMPLE = os.path.join("examples", "taichi_volume_raycaster.py")


def load_example():
    """The reference's TF-optimisation demo (examples/taichi_volume_raycaster.py: its own copy of the kernels plus `backward()`,
    :425-447, and the momentum step `apply_grad`, :375-381) on the interpreter.  Plotting and torchvtk imports are stubbed; jitter comes
    from ti_shim.set_random_source (the file is executed unmodified)."""
    import types
    import warnings
    from differender_b200.utils import tex_from_pts
    root = os.path.dirname(os.path.dirname(tp.find_reference()))
    path = os.path.join(root, EXAMPLE)
    stubs = {n: types.ModuleType(n) for n in ("matplotlib", "matplotlib.pyplot", "torchvtk", "torchvtk.rendering", "torchvtk.utils")}
    stubs["matplotlib"].pyplot = stubs["matplotlib.pyplot"]
    stubs["torchvtk.rendering"].plot_tf = stubs["torchvtk.rendering"].plot_tfs = None
    stubs["torchvtk.utils"].tex_from_pts, stubs["torchvtk.utils"].TFGenerator = tex_from_pts, None
    stubs["torchvtk"].rendering, stubs["torchvtk"].utils = stubs["torchvtk.rendering"], stubs["torchvtk.utils"]
    sys.path.insert(0, root)                                   # `from differender.utils import get_tf`
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            return ti_shim.load_reference(open(path).read(), path, extra_modules=stubs)
    finally:
        sys.path.remove(root)
        for k in [k for k in sys.modules if k == "differender" or k.startswith("differender.")]:
            del sys.modules[k]


# the eye sits inside the box: every ray has tens of samples, none has exactly one (whose position is 0/0 in the reference)
LOOP = dict(name="tf loop 12^3 black to tf1 R=32 16x16 sr 0.7", shape=(12, 12, 12), R=32, res=(16, 16), M=64, bw_sr=0.7, fw_sr=2.0,
            cam_pos=(0.5, 0.3, -0.8), iters=3, lr=0.1, mom=0.9, clip=0.1, decay=0.99)


def main_loop():
    """C2 in miniature, every arithmetic step executed by reference code on the interpreter: the LIBRARY's kernels
    (differender/volume_raycaster.py, the path under test) driven in the order of the demo's `backward()`
    (examples/taichi_volume_raycaster.py:425-447: clear, compute_entry_exit with jitter, raycast, get_final_image,
    torch.nn.functional.mse_loss against the reference image, get_final_image.grad, raycast.grad) and the DEMO's own momentum
    step kernel `apply_grad` (:375-381) with its learning-rate decay (:596-602), from the `black` TF towards a target rendered with
    `tf1` by the library's nondiff march.  (The demo file carries an older copy of the march kernels -- no min(1, .) on the Phong
    factor, no min(1, rgba) on the nondiff image -- so its kernels are not the ones pinned here; only its optimiser step is.)"""
    import torch
    from differender_b200.synthetic import make_jitter, make_tf, make_volume
    from oracle import cpu_oracle as co
    c = LOOP
    lib = tp._load_reference_module("shim")
    demo = load_example()
    ti_shim.reset()
    vol = make_volume(c["shape"]).numpy()[0]
    cam = np.asarray(c["cam_pos"], np.float32)
    w, h = c["res"]
    D, Hv, Wv = vol.shape
    tf_target, tf0 = make_tf("tf1", c["R"]).numpy(), make_tf("black", c["R"]).numpy()
    vr = lib.VolumeRaycaster((Wv, D, Hv), c["res"], max_samples=c["M"], tf_resolution=c["R"])
    vr.set_cam_pos(torch.tensor(cam))
    vr.set_volume(torch.tensor(vol).permute(2, 0, 1).contiguous())
    vr.set_tf_tex(torch.tensor(tf_target).permute(1, 0).contiguous())
    vr.clear_framebuffer()                                                     # Raycaster.raycast_nondiff (:514-520)
    vr.compute_entry_exit(c["fw_sr"], 0)
    vr.raycast_nondiff(c["fw_sr"])
    vr.get_final_image_nondiff()
    target_raw = vr.output_rgba.to_torch().clone()
    opt = demo.VolumeRaycaster(volume_resolution=(4, 4, 4), render_resolution=(16, 16), max_samples=1, tf_resolution=c["R"])   # apply_grad only
    tf = tf0.copy()
    lr, tfs, grads, losses, jits = c["lr"], [], [], [], []
    for k in range(c["iters"]):
        jit = make_jitter(1, h, w, seed=900 + k)[0].numpy()
        vr.set_tf_tex(torch.tensor(tf).permute(1, 0).contiguous())
        tp.set_jitter(vr, jit)
        vr.clear_framebuffer(); vr.clear_grad()
        vr.compute_entry_exit(c["bw_sr"], 1)
        vr.raycast(c["bw_sr"])
        vr.get_final_image()
        out = vr.output_rgba.to_torch().requires_grad_(True)
        loss = torch.nn.functional.mse_loss(out, target_raw)
        loss.backward()
        vr.output_rgba.grad.from_torch(out.grad)
        vr.get_final_image.grad()
        vr.raycast.grad(c["bw_sr"])
        g = vr.tf_tex.grad.to_torch().numpy()                                   # (R, 4) fp32, as the demo's apply_grad reads it
        assert np.isfinite(g).all() and int((vr.sample_step_nums.to_torch() == 1).sum()) == 0
        opt.tf_tex.from_numpy(np.ascontiguousarray(tf.T)); opt.tf_tex.grad.from_numpy(g)
        opt.apply_grad(lr, c["mom"], c["clip"])
        lr *= c["decay"]
        tf = np.ascontiguousarray(opt.tf_tex.to_numpy().T)
        grads.append(g.T.copy()); losses.append(float(loss)); tfs.append(tf.copy()); jits.append(jit)
        print(f"iteration {k}: loss {losses[-1]:.6f}, |grad| {np.abs(g).max():.3e}, tf alpha max {tf[3].max():.4f}")
    out_dir = os.path.join(HERE, "shim_loop")
    os.makedirs(out_dir, exist_ok=True)
    path = os.path.join(out_dir, "l0_tf_loop.npz")
    np.savez_compressed(path, name=c["name"], volume=vol, cam=cam, tf_init=tf0, target=co._raw_to_image(target_raw.numpy()), jitter=np.stack(jits),
                        output_shape=np.array(c["res"]), max_samples=c["M"], bw_sampling_rate=c["bw_sr"], fw_sampling_rate=c["fw_sr"],
                        tf_target=tf_target, lr=c["lr"], momentum=c["mom"], clip=c["clip"], lr_decay=c["decay"],
                        tf_after=np.stack(tfs), grad_tf=np.stack(grads), loss=np.array(losses))
    print(f"{path}: {os.path.getsize(path)} bytes")


if __name__ == "__main__":
    if "--loop-only" in sys.argv:
        main_loop()
    else:
        if "--api-only" not in sys.argv:
            main()
        main_api()
        main_loop()
