#!/usr/bin/env python
"""oracle/taichi_probe.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Opportunistic parity probe against the REAL reference.

The reference's arithmetic is Taichi's JIT (taichi + taichi_glsl, both unpinned and not installable in this image), so the
oracle (oracle/cpu_ref.c) is unpinned against the real compiler's ROUNDING.  The first box on which `import taichi` succeeds pins it with this script:

  * it loads the reference's own source, differender/volume_raycaster.py, from $DIFFERENDER_REFERENCE, /root/reference or
    baseline/_ref, UNMODIFIED except for a three-line patch (`PATCH` below) that replaces the `ti.random` jitter (:255) by a
    jitter FIELD so that the reference and the oracle march identical sample positions (north_star: "identical inputs and
    jitter") -- the `--shim` engine needs no patch at all: the interpreter's ti.random() returns the supplied jitter;
  * it bypasses `Raycaster.__init__` (which hard-codes `ti.init(arch=ti.cuda)`, :486) and drives the reference's
    `VolumeRaycaster` (:56-389) by hand after `ti.init(arch=ti.cpu, default_fp=ti.f32)` with exactly the call sequence of
    `RaycastFunction.forward / backward` for one item (:431-438, :467-476);
  * it compares image, sample counts and both gradients with the default oracle AND with every rounding variant of
    oracle/cpu_oracle.py VARIANTS, and reports which build the real Taichi arithmetic agrees with.

Without Taichi the same driver can run the reference source on oracle/ti_shim.py, a strict-IEEE-fp32 interpreter of the Taichi
subset the reference uses (`--shim`, `probe(engine="shim")`): that pins the oracle's COMPOSITION against the reference's own source
(the `source_order` build must match it bit for bit) but not the real compiler's rounding -- see oracle/ti_shim.py.

When Taichi or the reference source is missing it says so and changes nothing:

    python oracle/taichi_probe.py            # prints the report (or "Taichi unavailable ...")
    python oracle/taichi_probe.py --shim     # the reference source on the fp32 interpreter vs every oracle build (CPU, ~1 min)
    from oracle import taichi_probe; taichi_probe.status()   # one-line dict for bench.py / smoke()
"""
import importlib.util
import os
import sys
import types

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)

# (old, new) source substitutions; each `old` must occur exactly once in the reference file
PATCH = (
    # 1. a field for the supplied jitter, declared next to the other per-pixel fields (:96)
    ("        self.cam_pos = ti.Vector.field(3, dtype=ti.f32)\n",
     "        self.cam_pos = ti.Vector.field(3, dtype=ti.f32)\n        self.jitter_field = ti.field(ti.f32)\n"),
    # 2. placed in the same 8x8-tiled layout as entry / exit (:108-109)
    ("        ti.root.place(self.cam_pos)\n",
     "        ti.root.place(self.cam_pos)\n        ti.root.dense(ti.ij, render_resolution).dense(ti.ij, (8, 8)).place(self.jitter_field)\n"),
    # 3. read instead of ti.random (:255)
    ("ti.random(dtype=float)", "self.jitter_field[i, j]"),
)
REL = os.path.join("differender", "volume_raycaster.py")


def find_reference():
    """Path of the reference's differender/volume_raycaster.py, or None."""
    roots = [os.environ.get("DIFFERENDER_REFERENCE"), "/root/reference", os.path.join(_ROOT, "baseline", "_ref")]
    for r in roots:
        if r and os.path.isfile(os.path.join(r, REL)):
            return os.path.join(r, REL)
    return None


def patched_source(path):
    """The reference source with PATCH applied (raises if the file is not the one the patch was written for)."""
    src = open(path).read()
    for old, new in PATCH:
        if src.count(old) != 1:
            raise RuntimeError(f"taichi_probe: patch anchor {old!r} occurs {src.count(old)} times in {path} (expected 1)")
        src = src.replace(old, new)
    return src


def taichi_status():
    """(ok, message): whether taichi and taichi_glsl import."""
    for mod in ("taichi", "taichi_glsl"):
        if importlib.util.find_spec(mod) is None:
            return False, f"Taichi unavailable: no module named '{mod}' (not installable in this image: no network)"
    try:
        import taichi  # noqa: F401
        import taichi_glsl  # noqa: F401
    except Exception as e:  # a broken install is the same as none
        return False, f"Taichi unavailable: import failed ({type(e).__name__}: {e})"
    return True, "taichi importable"


def status():
    """One-line summary for bench.py / smoke(): never raises, never imports the product."""
    ok, msg = taichi_status()
    ref = find_reference()
    if not ok:
        return {"taichi": False, "pinned": False, "source_pinned": True,
                "note": msg + "; CPU restatement (oracle/cpu_ref.c) used: pinned against the reference's source on the fp32 interpreter "
                "(oracle/ti_shim.py, tests/test_shim_pin.py), unpinned against the real Taichi compiler's rounding"}
    if ref is None:
        return {"taichi": True, "pinned": False, "note": "taichi importable but the reference source (differender/volume_raycaster.py) "
                "was not found; set DIFFERENDER_REFERENCE and run oracle/taichi_probe.py"}
    try:
        rep = probe(cases=[CASES[0]])
        worst = rep["default"]
        return {"taichi": True, "pinned": True, "note": f"real Taichi (ti.cpu) vs default oracle on {CASES[0]['name']}: RGBA max-abs "
                f"{worst['max_abs']:.2e}, gvol rel-L2 {worst['gvol']:.2e}, gtf rel-L2 {worst['gtf']:.2e}; best-matching build: {rep['best']}"}
    except Exception as e:
        return {"taichi": True, "pinned": False, "note": f"taichi probe failed: {type(e).__name__}: {e}"}


# small cases: dims multiples of 4 (volume, :97) / 8 (image, :98); max_samples >= max n; tape = 16*W*H*M bytes x 2
CASES = [
    dict(name="64^3 tf1 64x48", n=64, tf="tf1", res=(64, 48), M=512, sr=1.0),
    dict(name="128^3 tf1 96x64", n=128, tf="tf1", res=(96, 64), M=1024, sr=1.0),
    dict(name="64^3 tf1 64x48 sr=2", n=64, tf="tf1", res=(64, 48), M=1024, sr=2.0),
]


def _load_reference_module(engine="taichi"):
    path = find_reference()
    if engine == "shim":
        import warnings
        from oracle import ti_shim
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", FutureWarning)       # torch.cuda.amp decorators of the reference's autograd wrapper
            return ti_shim.load_reference(open(path).read(), path)     # UNMODIFIED: the interpreter serves ti.random from the supplied jitter
    import taichi as ti
    ti.init(arch=ti.cpu, default_fp=ti.f32)                      # the library hard-codes ti.cuda (:486); the probe runs ti.cpu
    mod = types.ModuleType("differender_reference_patched")
    exec(compile(patched_source(path), path + " [+taichi_probe.PATCH]", "exec"), mod.__dict__)
    return mod


def set_jitter(vr, jit, copies=1):
    """Makes the reference march the supplied jitter image ((H, W), image orientation).  Real Taichi: the jitter field of PATCH.  The
    interpreter runs the UNMODIFIED source and serves its ti.random() calls -- one per pixel, in struct-for order = the raw (w, h)
    layout -- from the same numbers (`copies` forwards' worth: the batched backward re-runs compute_entry_exit, :456)."""
    import torch
    from oracle.cpu_oracle import _jitter_raw
    raw = _jitter_raw(jit)
    if hasattr(vr, "jitter_field"):
        vr.jitter_field.from_torch(torch.tensor(raw))
    else:
        from oracle import ti_shim
        ti_shim.set_random_source(np.tile(raw.reshape(-1), copies))


def run_reference(mod, vol, tf, cam, jit, grad_img, res, M, sr, nondiff=False, fov=30.0, near=0.1):
    """One item through the reference's VolumeRaycaster (real Taichi or the shim), following RaycastFunction.forward/backward
    (:431-438, :467-476) or Raycaster.raycast_nondiff (:514-520).  vol (D,H,W) fp32; tf (4,R); cam (3,); jit (H,W) or None and
    grad_img (4,H,W) in image orientation.  The gradient seed of rays with n <= 1 is zeroed (their sample position
    is 0/0 in the reference).  Returns a dict: image (4,H,W), K, n (H,W), grad_image as used, and -- unless nondiff -- gvol (D,H,W),
    gtf (4,R) after the reference's nan_to_num, plus the masks of the entries that were NaN before it (H4)."""
    import torch
    from oracle.cpu_oracle import _image_to_raw, _jitter_raw, _raw_to_image
    D, Hv, Wv = vol.shape
    vr = mod.VolumeRaycaster((Wv, D, Hv), res, max_samples=M, tf_resolution=tf.shape[1], fov=fov, nearfar=(near, 100.0))   # Taichi order (X,Y,Z) = torch (W,D,H) :481
    vr.set_cam_pos(torch.tensor(cam))
    vr.set_volume(torch.tensor(vol).permute(2, 0, 1).contiguous())                              # _determine_batch :571
    vr.set_tf_tex(torch.tensor(tf).permute(1, 0).contiguous())
    if jit is not None:
        set_jitter(vr, jit)
    vr.clear_framebuffer()
    vr.compute_entry_exit(sr, 0 if (jit is None or nondiff) else 1)
    n = _raw_to_image(vr.sample_step_nums.to_torch().numpy()[..., None])[0]
    if nondiff:
        vr.raycast_nondiff(sr)
        vr.get_final_image_nondiff()
        return dict(image=_raw_to_image(vr.output_rgba.to_torch().numpy()), n=n)
    vr.raycast(sr)
    vr.get_final_image()
    raw = vr.output_rgba.to_torch().numpy()
    K = vr.valid_sample_step_count.to_torch().numpy() - 1                                       # :303, :367
    go = np.where(n[None] > 1, grad_img, 0).astype(np.float32)
    vr.clear_grad()
    vr.output_rgba.grad.from_torch(torch.tensor(_image_to_raw(go)))
    vr.get_final_image.grad()
    vr.raycast.grad(sr)
    gv_raw, gt_raw = vr.volume.grad.to_torch(), vr.tf_tex.grad.to_torch()
    if hasattr(vr.volume.grad, "to_numpy64"):                                                   # the shim accumulates adjoints in float64
        gv64, gt64 = torch.from_numpy(vr.volume.grad.to_numpy64()), torch.from_numpy(vr.tf_tex.grad.to_numpy64())
    else:
        gv64, gt64 = gv_raw.double(), gt_raw.double()
    return dict(image=_raw_to_image(raw), K=_raw_to_image(K[..., None])[0], n=n, grad_image=go,
                gvol=torch.nan_to_num(gv64).permute(1, 2, 0).contiguous().numpy(),             # (X,Y,Z) -> torch (D,H,W); :474-475
                gtf=torch.nan_to_num(gt64).permute(1, 0).contiguous().numpy(),
                gvol_nan=torch.isnan(gv_raw).permute(1, 2, 0).contiguous().numpy(), gtf_nan=torch.isnan(gt_raw).permute(1, 0).contiguous().numpy())


def run_taichi(mod, vol, tf, cam, jit, grad_img, res, M, sr):
    """run_reference as the (image, K, n, gvol, gtf) tuple _cmp takes."""
    r = run_reference(mod, vol, tf, cam, jit, grad_img, res, M, sr)
    return r["image"], r["K"], r["n"], r["gvol"], r["gtf"]


def _cmp(ref, ora):
    """ref: run_reference's dict; ora: (image, K, n, gvol, gtf) of one oracle build.  Gradients are compared outside the entries the
    reference poisoned with NaN and then zeroed (the oracle keeps their finite contributions by design)."""
    i1, K1, n1, gv1, gt1 = ora
    i0, K0, n0 = ref["image"], ref["K"], ref["n"]
    ok = (K0 == K1) & (n0 > 1)                                   # n == 1 rays are 0/0 in the reference
    d = np.abs(np.asarray(i1, np.float64) - i0).max(axis=0)
    bits = (np.ascontiguousarray(i1, np.float32).view(np.uint32) != np.ascontiguousarray(i0, np.float32).view(np.uint32)).any(axis=0)

    def rel(a, b, nan):
        a, b = np.where(nan, 0, np.asarray(a, np.float64)), np.where(nan, 0, np.asarray(b, np.float64))
        return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))
    return dict(n_diff=int((n0 != n1).sum()), K_diff=int((K0 != K1).sum()), max_abs=float(d[ok].max()) if ok.any() else float("nan"),
                within=float((d[ok] <= 1e-4).mean()) if ok.any() else float("nan"), bit_diff=int(bits[ok].sum()),
                gvol=rel(gv1, ref["gvol"], ref["gvol_nan"]), gtf=rel(gt1, ref["gtf"], ref["gtf_nan"]),
                poisoned=int(ref["gvol_nan"].sum()) + int(ref["gtf_nan"].sum()))


# the interpreter costs ~1 ms per sample: a few hundred rays through a small volume per case
SHIM_CASES = [
    dict(name="12^3 tf1 R=32 16x8", shape=(12, 12, 12), tf="tf1", R=32, res=(16, 8), M=64, sr=1.0, jitter=True, cam=1),
    dict(name="16x12x20 rand R=16 16x16 sr=0.7", shape=(16, 12, 20), tf="rand", R=16, res=(16, 16), M=64, sr=0.7, jitter=True, cam=5),
    dict(name="12x16x8 tf3 R=24 8x16 sr=2 no jitter", shape=(12, 16, 8), tf="tf3", R=24, res=(8, 16), M=128, sr=2.0, jitter=False, cam=11),
    # the TF resolution of the benchmark configs, more rays (about a minute and ~1 GB of tape on the interpreter)
    dict(name="24^3 tf1 R=128 32x24", shape=(24, 24, 24), tf="tf1", R=128, res=(32, 24), M=128, sr=1.0, jitter=True, cam=7),
    # the camera INSIDE the box (entry distance negative: the reference marches from behind the eye) and another frustum
    dict(name="12^3 tf2 R=32 16x8 camera inside fov 45 near 0.5", shape=(12, 12, 12), tf="tf2", R=32, res=(16, 8), M=128, sr=1.0, jitter=True,
         cam_pos=(0.3, 0.2, -0.55), fov=45.0, near=0.5),
    # a block of exactly constant voxels: zero local gradient, normalized() = 0/0
    dict(name="12^3 gray R=8 8x8 with a flat block", shape=(12, 12, 12), tf="gray", R=8, res=(8, 8), M=64, sr=1.0, jitter=True, cam=3, flat=True),
    # Raycaster.raycast_nondiff (:490-523): forward only, alpha gate 1e-3, no shading clamp, min(1, rgba) at the end, sr = 4
    dict(name="12^3 tf1 R=32 16x8 nondiff sr=4", shape=(12, 12, 12), tf="tf1", R=32, res=(16, 8), M=256, sr=4.0, jitter=False, cam=1, nondiff=True),
]


def case_inputs(c):
    """Seeded inputs of one case (numpy): volume (D,H,W), tf (4,R), cam (3,), jitter (H,W) or None, grad_image (4,H,W)."""
    sys.path.insert(0, _ROOT)
    import torch
    from differender_b200.synthetic import make_cameras, make_jitter, make_tf, make_volume
    shape = c.get("shape", (c.get("n"),) * 3)
    vol = make_volume(shape).numpy()[0]
    if c.get("flat"):
        vol = vol.copy()
        vol[3:9, 2:8, 4:10] = 0.4375
    R = c.get("R", 128)
    if c["tf"] == "rand":
        tf = torch.rand(4, R, generator=torch.Generator().manual_seed(11))
        tf[3] *= 0.25
        tf = tf.numpy()
    else:
        tf = make_tf(c["tf"], R).numpy()
    cam = np.asarray(c["cam_pos"], np.float32) if "cam_pos" in c else make_cameras(16)[c.get("cam", 1)].numpy()
    h, w = c["res"][1], c["res"][0]
    jit = make_jitter(1, h, w)[0].numpy() if c.get("jitter", True) else None
    go = np.random.default_rng(7).standard_normal((4, h, w)).astype(np.float32)
    return vol, tf, cam, jit, go


def probe(cases=None, out=None, engine="taichi"):
    """Runs the cases through the reference (real Taichi, or its source on the fp32 interpreter with engine="shim") and through every
    oracle build; returns {"default": worst-case row, "best": name, "rows": {case: {build: row}}, ...}."""
    if engine == "taichi":
        ok, msg = taichi_status()
        if not ok:
            raise RuntimeError(msg)
    elif find_reference() is None:
        raise RuntimeError("the reference source (differender/volume_raycaster.py) was not found")
    sys.path.insert(0, _ROOT)
    from oracle import cpu_oracle as co
    co.build()
    mod = _load_reference_module(engine)
    what = "real Taichi ti.cpu" if engine == "taichi" else "the reference SOURCE on oracle/ti_shim.py (strict IEEE fp32, source order)"
    lines, score, rows = [], {}, {}
    worst_default = dict(max_abs=0.0, gvol=0.0, gtf=0.0)
    for c in (cases or (CASES if engine == "taichi" else SHIM_CASES)):
        vol, tf, cam, jit, go = case_inputs(c)
        fk = dict(fov=c.get("fov", 30.0), near=c.get("near", 0.1))
        if engine == "shim":
            from oracle import ti_shim
            ti_shim.reset()
        if c.get("nondiff"):
            ref = run_reference(mod, vol, tf, cam, None, None, c["res"], c["M"], c["sr"], nondiff=True, **fk)
            lines.append(f"## {c['name']}  ({what} vs oracle builds, forward only)")
            lines.append(f"{'oracle build':<20}{'n_diff':>8}{'px bits!=':>10}{'max_abs':>11}")
            rows[c["name"]] = {}
            for name in [None] + list(co.VARIANTS):
                img, K, n = co.forward(vol, tf, cam, c["res"], return_counts=True, sampling_rate=c["sr"], max_samples=c["M"], nondiff=True, variant=name, **fk)
                ok = ref["n"] > 1
                bits = (np.ascontiguousarray(img, np.float32).view(np.uint32) != ref["image"].view(np.uint32)).any(axis=0)
                r = dict(n_diff=int((n != ref["n"]).sum()), bit_diff=int(bits[ok].sum()), max_abs=float(np.abs(img - ref["image"]).max(axis=0)[ok].max()))
                rows[c["name"]][name or "default"] = r
                score[name or "default"] = max(score.get(name or "default", 0.0), r["max_abs"] / 1e-4)
                lines.append(f"{name or 'default':<20}{r['n_diff']:>8}{r['bit_diff']:>10}{r['max_abs']:>11.2e}")
            lines.append("")
            continue
        ref = run_reference(mod, vol, tf, cam, jit, go, c["res"], c["M"], c["sr"], **fk)
        lines.append(f"## {c['name']}  ({what} vs oracle builds; rays with n <= 1 masked, SURVEY H3: {int((ref['n'] == 1).sum())} here; "
                     f"{int(ref['gvol_nan'].sum())} voxels / {int(ref['gtf_nan'].sum())} TF entries NaN-poisoned in the reference, H4)")
        lines.append(f"{'oracle build':<20}{'n_diff':>8}{'K_diff':>8}{'px bits!=':>10}{'max_abs':>11}{'within 1e-4':>13}{'gvol relL2':>12}{'gtf relL2':>12}")
        rows[c["name"]] = {}
        for name in [None] + list(co.VARIANTS):
            kw = dict(sampling_rate=c["sr"], max_samples=c["M"], jitter=jit, variant=name, **fk)
            img, K, n = co.forward(vol, tf, cam, c["res"], return_counts=True, **kw)
            gv, gt = co.backward(vol, tf, cam, ref["grad_image"], c["res"], **kw)
            r = _cmp(ref, (img, K, n, gv, gt))
            label = name or "default"
            rows[c["name"]][label] = r
            score[label] = max(score.get(label, 0.0), r["max_abs"] / 1e-4, r["gvol"] / 1e-3, r["gtf"] / 1e-3)
            if name is None:
                for k in worst_default:
                    worst_default[k] = max(worst_default[k], r[k])
            lines.append(f"{label:<20}{r['n_diff']:>8}{r['K_diff']:>8}{r['bit_diff']:>10}{r['max_abs']:>11.2e}{100 * r['within']:>12.3f}%{r['gvol']:>12.2e}{r['gtf']:>12.2e}")
        lines.append("")
    best = min(score, key=score.get)
    lines.append(f"best-matching oracle build: {best} (worst tolerance fraction {score[best]:.3g}); default: {score['default']:.3g}  "
                 f"(< 1 means inside the north_star tolerances: RGBA 1e-4 max-abs, gradients 1e-3 rel-L2)")
    text = "\n".join(lines)
    if out:
        with open(out, "w") as f:
            f.write(text + "\n")
    return {"default": worst_default, "best": best, "score": score, "rows": rows, "text": text}


def main(argv=None):
    argv = sys.argv[1:] if argv is None else argv
    ref = find_reference()
    if "--shim" in argv:
        if ref is None:
            print("the reference source (differender/volume_raycaster.py) was not found: set DIFFERENDER_REFERENCE")
            return 1
        rep = probe(engine="shim", out=os.path.join(_ROOT, "profiles", "r02_shim_pin_report.txt"))
        print(rep["text"])
        return 0
    ok, msg = taichi_status()
    if not ok:
        print(msg + " -- CPU restatement (oracle/cpu_ref.c) used; parity unpinned against the real Taichi arithmetic "
              "(`--shim` runs the reference source on the fp32 interpreter instead)")
        if ref:
            patched_source(ref)
            print(f"(reference source found at {ref}; the 3-line jitter patch applies cleanly)")
        return 0
    if ref is None:
        print("taichi importable, but differender/volume_raycaster.py was not found: set DIFFERENDER_REFERENCE")
        return 1
    rep = probe(out=os.path.join(_ROOT, "profiles", "taichi_probe_report.txt"))
    print(rep["text"])
    return 0


if __name__ == "__main__":
    sys.exit(main())
