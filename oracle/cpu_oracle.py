"""ctypes front-end of oracle/cpu_ref.c (TEST INFRASTRUCTURE, not product code).

Mirrors the reference's L3 wrapper (differender/volume_raycaster.py:478-574): takes the volume in the
torch layout (D,H,W), the TF as (4,R), the camera as (3,), and returns images as (4,H,W) with the
reference's flip/permute (:538-548, 513, 523) done in numpy.  Jitter tensors are in the output-image
orientation (H,W): jit(i,j) = J[H-1-j, i].
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {}


class OraDesc(ctypes.Structure):
    _fields_ = [("X", ctypes.c_int), ("Y", ctypes.c_int), ("Z", ctypes.c_int),
                ("W", ctypes.c_int), ("H", ctypes.c_int), ("R", ctypes.c_int), ("M", ctypes.c_int),
                ("sr", ctypes.c_double), ("fov_deg", ctypes.c_double), ("near_", ctypes.c_double),
                ("nondiff", ctypes.c_int), ("has_jitter", ctypes.c_int)]


def build(force=False):
    """Compile the oracle with the recipe in oracle/Makefile (building the checker is not using it)."""
    outs = [os.path.join(_HERE, "build", n) for n in ("liboracle.so", "liboracle_fp64.so", "liboracle_fast.so")]
    # always through make: it compares the mtimes of cpu_ref.c and the libraries (a no-op when they are current)
    subprocess.check_call(["make", "-C", _HERE] + (["-B"] if force else []), stdout=subprocess.DEVNULL)
    return outs


# Rounding-envelope variants of the strict fp32 oracle (cpu_ref.c header): name -> -D switches.  Diagnostic only
# (tools/rounding_envelope.py); no parity test uses them.
VARIANTS = {
    "mix_unfused": ["ORA_MIX_UNFUSED"],
    "mix_fma_b": ["ORA_MIX_FMA_B"],
    "div_true": ["ORA_DIV_TRUE"],
    "div_approx": ["ORA_DIV_APPROX"],
    "pow_fast": ["ORA_POW_FAST"],
    "tan_f32": ["ORA_TAN_F32"],
    "pos_unfused": ["ORA_POS_UNFUSED"],
    "composite_unfused": ["ORA_COMPOSITE_UNFUSED"],
    "normalize_div": ["ORA_NORMALIZE_DIV"],
    # one IEEE-rounded operation per SOURCE operator, in source order: what oracle/ti_shim.py computes when it interprets the
    # reference's own source -- this build must agree with it bit for bit (tests/test_shim_pin.py)
    "source_order": ["ORA_MIX_UNFUSED", "ORA_POS_UNFUSED", "ORA_COMPOSITE_UNFUSED", "ORA_DIV_TRUE"],
    # no contraction anywhere + true division: what a target without FMA contraction (or fast_math=False) computes
    "strict_ieee": ["ORA_MIX_UNFUSED", "ORA_POS_UNFUSED", "ORA_COMPOSITE_UNFUSED", "ORA_DIV_TRUE", "ORA_NORMALIZE_DIV"],
    # every fast-math liberty at once, the other way
    "all_fast": ["ORA_MIX_FMA_B", "ORA_DIV_APPROX", "ORA_POW_FAST", "ORA_TAN_F32"],
}


def build_variant(name):
    """Compile one rounding variant of the strict fp32 oracle into oracle/build/variants/ (same flags as liboracle.so)."""
    out = os.path.join(_HERE, "build", "variants", f"liboracle_{name}.so")
    src = os.path.join(_HERE, "cpu_ref.c")
    if not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.check_call(["/usr/bin/gcc", "-std=c11", "-ffp-contract=off", "-fno-fast-math", "-fopenmp", "-fPIC", "-shared",
                               "-O2", "-mfma"] + ["-D" + d for d in VARIANTS[name]] + ["-o", out, src, "-lm"])
    return out


def _lib(fast=False, fp64=False, variant=None):
    key = ("variant:" + variant) if variant else ("fp64" if fp64 else ("fast" if fast else "strict"))
    if key not in _LIBS:
        if variant:
            path = build_variant(variant)
        else:
            path = os.path.join(_HERE, "build", {"fp64": "liboracle_fp64.so", "fast": "liboracle_fast.so",
                                                 "strict": "liboracle.so"}[key])
        if not os.path.exists(path):
            build()
        lib = ctypes.CDLL(path)
        fp = ctypes.POINTER(ctypes.c_double if fp64 else ctypes.c_float)
        dp = ctypes.POINTER(ctypes.c_double)
        ip = ctypes.POINTER(ctypes.c_int32)
        lib.ora_forward.argtypes = [ctypes.POINTER(OraDesc), fp, fp, fp, fp, fp, ip, ip]
        lib.ora_forward.restype = None
        lib.ora_backward.argtypes = [ctypes.POINTER(OraDesc), fp, fp, fp, fp, fp, dp, dp, ctypes.c_int]
        lib.ora_backward.restype = None
        lib.ora_num_threads.restype = ctypes.c_int
        lib.ora_set_num_threads.argtypes = [ctypes.c_int]
        _LIBS[key] = lib
    return _LIBS[key]


def num_threads(fast=False):
    return _lib(fast).ora_num_threads()


def set_num_threads(n, fast=False):
    _lib(fast).ora_set_num_threads(int(n))


def _f32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def _real(a, fp64):
    """fp32 library: inputs rounded to fp32.  fp64 library: inputs widened to (or kept as) double."""
    if a is None:
        return None
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64)) if fp64 else _f32(a)


def _ptr(a, ty=None):
    if a is None:
        return None
    if ty is None:
        ty = ctypes.c_double if a.dtype == np.float64 else ctypes.c_float
    return a.ctypes.data_as(ctypes.POINTER(ty))


def _desc(vol, tf, out_shape, sr, max_samples, fov, near, nondiff, has_jitter):
    D, Hv, Wv = vol.shape[-3:]
    w, h = out_shape
    return OraDesc(X=Wv, Y=D, Z=Hv, W=w, H=h, R=tf.shape[-1], M=max_samples, sr=float(sr),
                   fov_deg=float(fov), near_=float(near), nondiff=int(nondiff), has_jitter=int(has_jitter))


def _raw_to_image(raw):
    """(W,H,C) raw -> (C,H,W) flipped, as volume_raycaster.py:543-548."""
    return np.ascontiguousarray(np.flip(raw, axis=1).transpose(2, 1, 0))


def _image_to_raw(img):
    """(C,H,W) image orientation -> (W,H,C) raw (inverse of the above)."""
    return np.ascontiguousarray(np.flip(img.transpose(2, 1, 0), axis=1))


def _jitter_raw(jitter):
    if jitter is None:
        return None
    return np.ascontiguousarray(_image_to_raw(np.asarray(jitter)[None])[..., 0])


def forward(volume, tf, look_from, output_shape, sampling_rate=1.0, max_samples=512, fov=30.0, near=0.1,
            jitter=None, nondiff=False, fast=False, fp64=False, return_counts=False, variant=None):
    """One view.  volume (D,H,W) or (1,D,H,W); tf (4,R); look_from (3,).  Returns (4,H,W) float32
    (float64 with fp64=True) [and K (H,W), n (H,W) int32 in image orientation]."""
    vol = _real(volume, fp64).reshape(np.asarray(volume).shape[-3:])
    tfa = _real(tf, fp64)
    tf_r4 = np.ascontiguousarray(tfa.T)                     # tf.permute(1,0)  (:571)
    cam = _real(look_from, fp64).reshape(3)
    w, h = output_shape
    d = _desc(vol, tfa, output_shape, sampling_rate, max_samples, fov, near, nondiff, jitter is not None)
    jr = _real(_jitter_raw(jitter), fp64)
    out = np.zeros((w, h, 4), vol.dtype)
    K = np.zeros((w, h), np.int32)
    n = np.zeros((w, h), np.int32)
    _lib(fast, fp64, variant).ora_forward(ctypes.byref(d), _ptr(vol), _ptr(tf_r4), _ptr(cam), _ptr(jr), _ptr(out),
                           _ptr(K, ctypes.c_int32), _ptr(n, ctypes.c_int32))
    img = _raw_to_image(out)
    if return_counts:
        return img, _raw_to_image(K[..., None])[0], _raw_to_image(n[..., None])[0]
    return img


def backward(volume, tf, look_from, grad_image, output_shape, sampling_rate=1.0, max_samples=512, fov=30.0,
             near=0.1, jitter=None, want_vol=True, want_tf=True, fast=False, fp64=False, variant=None):
    """One view.  grad_image (4,H,W).  Returns (grad_volume (D,H,W) float64 or None, grad_tf (4,R) float64 or None)."""
    vol = _real(volume, fp64).reshape(np.asarray(volume).shape[-3:])
    tfa = _real(tf, fp64)
    tf_r4 = np.ascontiguousarray(tfa.T)
    cam = _real(look_from, fp64).reshape(3)
    d = _desc(vol, tfa, output_shape, sampling_rate, max_samples, fov, near, False, jitter is not None)
    jr = _real(_jitter_raw(jitter), fp64)
    go = _real(_image_to_raw(np.asarray(grad_image)), fp64)
    gvol = np.zeros(vol.shape, np.float64)
    gtf = np.zeros(tf_r4.shape, np.float64)
    flags = (1 if want_vol else 0) | (2 if want_tf else 0)
    _lib(fast, fp64, variant).ora_backward(ctypes.byref(d), _ptr(vol), _ptr(tf_r4), _ptr(cam), _ptr(jr), _ptr(go),
                            _ptr(gvol, ctypes.c_double), _ptr(gtf, ctypes.c_double), flags)
    return (gvol if want_vol else None), (np.ascontiguousarray(gtf.T) if want_tf else None)
