#!/usr/bin/env python
"""oracle/taichi_probe.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Opportunistic parity probe against the REAL reference.

The reference's arithmetic is Taichi's JIT (taichi + taichi_glsl, both unpinned and not installable in this image), so the
oracle (oracle/cpu_ref.c) is "parity unpinned".  The first box on which `import taichi` succeeds pins it with this script:

  * it loads the reference's own source, differender/volume_raycaster.py, from $DIFFERENDER_REFERENCE, /root/reference or
    baseline/_ref, UNMODIFIED except for a three-line patch (`PATCH` below) that replaces the `ti.random` jitter (:255) by a
    jitter FIELD so that the reference and the oracle march identical sample positions (north_star: "identical inputs and
    jitter");
  * it bypasses `Raycaster.__init__` (which hard-codes `ti.init(arch=ti.cuda)`, :486) and drives the reference's
    `VolumeRaycaster` (:56-389) by hand after `ti.init(arch=ti.cpu, default_fp=ti.f32)` with exactly the call sequence of
    `RaycastFunction.forward / backward` for one item (:431-438, :467-476);
  * it compares image, sample counts and both gradients with the default oracle AND with every rounding variant of
    oracle/cpu_oracle.py VARIANTS, and reports which build the real Taichi arithmetic agrees with.

When Taichi or the reference source is missing it says so and changes nothing:

    python oracle/taichi_probe.py            # prints the report (or "Taichi unavailable ...")
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
        return {"taichi": False, "pinned": False, "note": msg + "; CPU restatement (oracle/cpu_ref.c) used, parity unpinned"}
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


def _load_reference_module():
    import taichi as ti
    ti.init(arch=ti.cpu, default_fp=ti.f32)                      # the library hard-codes ti.cuda (:486); the probe runs ti.cpu
    mod = types.ModuleType("differender_reference_patched")
    path = find_reference()
    exec(compile(patched_source(path), path + " [+taichi_probe.PATCH]", "exec"), mod.__dict__)
    return mod


def run_taichi(mod, vol, tf, cam, jit, grad_img, res, M, sr):
    """One item through the reference's VolumeRaycaster, following RaycastFunction.forward/backward (:431-438, :467-476).
    vol (D,H,W) fp32; tf (4,R); cam (3,); jit, grad_img in image orientation ((H,W), (4,H,W)).  Returns image (4,H,W), K (H,W),
    n (H,W), grad volume (D,H,W), grad tf (4,R)."""
    import torch
    from oracle.cpu_oracle import _image_to_raw, _jitter_raw, _raw_to_image
    D, Hv, Wv = vol.shape
    vr = mod.VolumeRaycaster((Wv, D, Hv), res, max_samples=M, tf_resolution=tf.shape[1])       # Taichi order (X,Y,Z) = torch (W,D,H) :481
    vr.set_cam_pos(torch.tensor(cam))
    vr.set_volume(torch.tensor(vol).permute(2, 0, 1).contiguous())                              # _determine_batch :571
    vr.set_tf_tex(torch.tensor(tf).permute(1, 0).contiguous())
    vr.jitter_field.from_torch(torch.tensor(_jitter_raw(jit)))
    vr.clear_framebuffer()
    vr.compute_entry_exit(sr, 1)
    vr.raycast(sr)
    vr.get_final_image()
    raw = vr.output_rgba.to_torch().numpy()
    K = vr.valid_sample_step_count.to_torch().numpy() - 1                                       # :303, :367
    n = vr.sample_step_nums.to_torch().numpy()
    vr.clear_grad()
    vr.output_rgba.grad.from_torch(torch.tensor(_image_to_raw(grad_img)))
    vr.get_final_image.grad()
    vr.raycast.grad(sr)
    gv = torch.nan_to_num(vr.volume.grad.to_torch()).permute(1, 2, 0).numpy()                   # (X,Y,Z) -> torch (D,H,W)
    gt = torch.nan_to_num(vr.tf_tex.grad.to_torch()).permute(1, 0).numpy()
    return _raw_to_image(raw), _raw_to_image(K[..., None])[0], _raw_to_image(n[..., None])[0], gv, gt


def _cmp(ti_out, ora_out):
    (i0, K0, n0, gv0, gt0), (i1, K1, n1, gv1, gt1) = ti_out, ora_out
    ok = (K0 == K1) & (n0 > 1)                                   # n == 1 rays are 0/0 in the reference
    d = np.abs(np.asarray(i1, np.float64) - i0).max(axis=0)
    rel = lambda a, b: float(np.linalg.norm(np.asarray(a, np.float64) - b) / max(np.linalg.norm(b), 1e-300))
    return dict(n_diff=int((n0 != n1).sum()), K_diff=int((K0 != K1).sum()), max_abs=float(d[ok].max()) if ok.any() else float("nan"),
                within=float((d[ok] <= 1e-4).mean()) if ok.any() else float("nan"), gvol=rel(gv1, gv0), gtf=rel(gt1, gt0))


def probe(cases=None, out=None):
    """Runs the cases through real Taichi and through every oracle build; returns {"default": worst-case row, "best": name, ...}."""
    ok, msg = taichi_status()
    if not ok:
        raise RuntimeError(msg)
    sys.path.insert(0, _ROOT)
    from differender_b200.synthetic import make_cameras, make_jitter, make_tf, make_volume
    from oracle import cpu_oracle as co
    co.build()
    mod = _load_reference_module()
    lines, score = [], {}
    worst_default = dict(max_abs=0.0, gvol=0.0, gtf=0.0)
    for c in (cases or CASES):
        vol = make_volume(c["n"]).numpy()[0]
        tf = make_tf(c["tf"], 128).numpy()
        cam = make_cameras(16)[1].numpy()
        h, w = c["res"][1], c["res"][0]
        jit = make_jitter(1, h, w)[0].numpy()
        go = np.random.default_rng(7).standard_normal((4, h, w)).astype(np.float32)
        t_out = run_taichi(mod, vol, tf, cam, jit, go, c["res"], c["M"], c["sr"])
        lines.append(f"## {c['name']}  (real Taichi ti.cpu vs oracle builds; rays with n <= 1 masked, SURVEY H3)")
        lines.append(f"{'oracle build':<20}{'n_diff':>8}{'K_diff':>8}{'max_abs':>11}{'within 1e-4':>13}{'gvol relL2':>12}{'gtf relL2':>12}")
        for name in [None] + list(co.VARIANTS):
            kw = dict(sampling_rate=c["sr"], max_samples=c["M"], jitter=jit, variant=name)
            img, K, n = co.forward(vol, tf, cam, c["res"], return_counts=True, **kw)
            gv, gt = co.backward(vol, tf, cam, go, c["res"], **kw)
            r = _cmp(t_out, (img, K, n, gv, gt))
            label = name or "default"
            score[label] = max(score.get(label, 0.0), r["max_abs"] / 1e-4, r["gvol"] / 1e-3, r["gtf"] / 1e-3)
            if name is None:
                for k in worst_default:
                    worst_default[k] = max(worst_default[k], r[k])
            lines.append(f"{label:<20}{r['n_diff']:>8}{r['K_diff']:>8}{r['max_abs']:>11.2e}{100 * r['within']:>12.3f}%{r['gvol']:>12.2e}{r['gtf']:>12.2e}")
        lines.append("")
    best = min(score, key=score.get)
    lines.append(f"best-matching oracle build: {best} (worst tolerance fraction {score[best]:.3f}); default: {score['default']:.3f}  "
                 f"(< 1 means inside the north_star tolerances: RGBA 1e-4 max-abs, gradients 1e-3 rel-L2)")
    text = "\n".join(lines)
    if out:
        with open(out, "w") as f:
            f.write(text + "\n")
    return {"default": worst_default, "best": best, "score": score, "text": text}


def main():
    ok, msg = taichi_status()
    ref = find_reference()
    if not ok:
        print(msg + " -- CPU restatement (oracle/cpu_ref.c) used; parity unpinned")
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
