"""ctypes front-end of tests/hostsim (TEST-ONLY host build of the kernels' per-ray arithmetic)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
# HOSTSIM_DEFINES="NAME ..." builds (into its own .so) a variant of the device math with tuning macros set, e.g. DR_DIRECT_TAPS
_DEFS = os.environ.get("HOSTSIM_DEFINES", "").split()
_SO = os.path.join(_HERE, "hostsim", "build", "libhostsim" + "".join("_" + d for d in _DEFS) + ".so")

# flag values of include/diffrender.h
F_NONDIFF, F_VOL, F_TF, F_JIT, F_IMG, F_TF4R, F_GENERIC, F_BRICK8, F_CELL8 = 1, 2, 4, 8, 16, 32, 64, 256, 2048


class DrDesc(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in ("X", "Y", "Z", "W", "H", "R", "M", "BS", "Bvol", "Btf", "vox_dtype")] + \
               [("flags", ctypes.c_uint32)] + \
               [(n, ctypes.c_float) for n in ("sr", "inv_sr", "near_", "near_w", "near_h")] + \
               [("scale", ctypes.c_float * 3)] + \
               [(n, ctypes.c_float) for n in ("vol_diag", "tf_len", "ambient", "diffuse", "specular", "ert", "delta",
                                              "alpha_skip")] + \
               [(n, ctypes.c_int32) for n in ("nbx", "nby", "nbz", "tap_generic")]


def build(force=False):
    src = os.path.join(_HERE, "hostsim", "hostsim.cpp")
    deps = [src] + [os.path.join(_ROOT, "differender_b200", "csrc", f) for f in ("dr_math.cuh", "dr_desc.h")] + \
           [os.path.join(_ROOT, "include", "diffrender.h")]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(p) > os.path.getmtime(_SO) for p in deps):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O2", "-mfma", "-ffp-contract=off", "-fno-fast-math",
                               "-fPIC", "-shared", "-x", "c++"] + ["-D" + d for d in _DEFS] + ["-I" + os.path.join(_ROOT, "include"),
                               "-I" + os.path.join(_ROOT, "differender_b200", "csrc"), "-o", _SO, src])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.sim_bricked_elems.restype = ctypes.c_size_t
    return _lib


def _p(a, ty=ctypes.c_float):
    return None if a is None else a.ctypes.data_as(ctypes.POINTER(ty))


def make_desc(vol_shape_dhw, output_shape, R, max_samples, flags, sr=1.0, fov=30.0, near=0.1):
    D, Hv, Wv = vol_shape_dhw
    d = DrDesc()
    rc = lib().sim_desc_init(ctypes.byref(d), Wv, D, Hv, output_shape[0], output_shape[1], R, max_samples,
                             ctypes.c_uint32(flags), ctypes.c_double(sr), ctypes.c_double(fov), ctypes.c_double(near))
    assert rc == 0
    return d


def _layout_data(L, d, vol, brick, cell):
    if brick:
        br = np.zeros(L.sim_bricked_elems(ctypes.byref(d)), np.float32)
        L.sim_brick(ctypes.byref(d), _p(vol), _p(br))
        return br
    if cell:
        ce = np.zeros(vol.size * 8, np.float32)
        L.sim_expand(ctypes.byref(d), _p(vol), _p(ce))
        return ce
    return vol


def macro_dims(d):
    """(ny, nz, nx, edge): macro-cells of the skip grid per axis and their edge length in cells."""
    out = (ctypes.c_int * 4)()
    lib().sim_macro_dims(ctypes.byref(d), out)
    return tuple(out)


def skip_grid(volume, tf, output_shape, sampling_rate=1.0, max_samples=512):
    """The macro-cell emptiness bytes dr_build_skip_grid would produce for (volume, tf): array [nby][nbz][nbx] of 0/1."""
    vol = np.ascontiguousarray(volume, np.float32).reshape(np.asarray(volume).shape[-3:])
    tf_r4 = np.ascontiguousarray(np.asarray(tf, np.float32).T)
    d = make_desc(vol.shape, output_shape, tf_r4.shape[0], max_samples, 0, sampling_rate)
    g = np.zeros(macro_dims(d)[:3], np.uint8)
    lib().sim_skip_grid(ctypes.byref(d), _p(vol), _p(tf_r4), _p(g, ctypes.c_ubyte))
    return g


def forward(volume, tf, cam, output_shape, sampling_rate=1.0, max_samples=512, jitter=None, nondiff=False, generic=False,
            brick=False, cell=False, skip=False, fov=30.0, near=0.1):
    vol = np.ascontiguousarray(volume, np.float32).reshape(np.asarray(volume).shape[-3:])
    tf_r4 = np.ascontiguousarray(np.asarray(tf, np.float32).T)
    flags = (F_NONDIFF if nondiff else 0) | (F_JIT if jitter is not None else 0) | (F_GENERIC if generic else 0) | F_IMG | \
            (F_BRICK8 if brick else 0) | (F_CELL8 if cell else 0)
    d = make_desc(vol.shape, output_shape, tf_r4.shape[0], max_samples, flags, sampling_rate, fov, near)
    L = lib()
    br = _layout_data(L, d, vol, brick, cell)
    w, h = output_shape
    out = np.zeros((4, h, w), np.float32); K = np.zeros((h, w), np.int32); Tp = np.zeros((h, w), np.float32)
    n = np.zeros((h, w), np.int32)
    cam = np.ascontiguousarray(cam, np.float32)
    jit = None if jitter is None else np.ascontiguousarray(jitter, np.float32)
    grid = None
    if skip:
        grid = np.zeros(macro_dims(d)[:3], np.uint8)
        L.sim_skip_grid(ctypes.byref(d), _p(vol), _p(tf_r4), _p(grid, ctypes.c_ubyte))
    L.sim_forward(ctypes.byref(d), _p(br), _p(tf_r4), _p(cam), _p(jit), _p(out), _p(K, ctypes.c_int32), _p(Tp),
                  _p(n, ctypes.c_int32), _p(grid, ctypes.c_ubyte))
    return out, K, Tp, n


def backward(volume, tf, cam, grad_image, output_shape, sampling_rate=1.0, max_samples=512, jitter=None,
             want_vol=True, want_tf=True, generic=False, brick=False, cell=False, skip=False, fov=30.0, near=0.1):
    vol = np.ascontiguousarray(volume, np.float32).reshape(np.asarray(volume).shape[-3:])
    tf_r4 = np.ascontiguousarray(np.asarray(tf, np.float32).T)
    out, K, Tp, _ = forward(volume, tf, cam, output_shape, sampling_rate, max_samples, jitter, False, generic, brick, cell, fov=fov, near=near)
    flags = (F_JIT if jitter is not None else 0) | (F_GENERIC if generic else 0) | F_IMG | (F_BRICK8 if brick else 0) | \
            (F_CELL8 if cell else 0) | (F_VOL if want_vol else 0) | (F_TF if want_tf else 0)
    d = make_desc(vol.shape, output_shape, tf_r4.shape[0], max_samples, flags, sampling_rate, fov, near)
    L = lib()
    br = _layout_data(L, d, vol, brick, cell)
    gbr = np.zeros(vol.size * 8, np.float32); gtf = np.zeros_like(tf_r4)
    cam = np.ascontiguousarray(cam, np.float32)
    jit = None if jitter is None else np.ascontiguousarray(jitter, np.float32)
    go = np.ascontiguousarray(grad_image, np.float32)
    grid = None
    if skip:                                               # the forward's skip grid: only the volume-only backward uses it
        grid = np.zeros(macro_dims(d)[:3], np.uint8)
        L.sim_skip_grid(ctypes.byref(d), _p(vol), _p(tf_r4), _p(grid, ctypes.c_ubyte))
    L.sim_backward(ctypes.byref(d), _p(br), _p(tf_r4), _p(cam), _p(jit), _p(go), _p(out), _p(K, ctypes.c_int32),
                   _p(Tp), _p(gbr), _p(gtf), _p(grid, ctypes.c_ubyte))
    gv = np.zeros_like(vol)
    L.sim_gather(ctypes.byref(d), _p(gbr), _p(gv))
    return gv, np.ascontiguousarray(gtf.T)


def gather(cells, vol_shape_dhw):
    """gather_voxel() over a cell-major gradient [D*H*W*8] -> [D,H,W] (host build of the device function; no nan_to_num)."""
    d = make_desc(vol_shape_dhw, (8, 8), 2, 1, 0)
    ce = np.ascontiguousarray(cells, np.float32).reshape(-1)
    out = np.zeros(vol_shape_dhw, np.float32)
    lib().sim_gather(ctypes.byref(d), _p(ce), _p(out))
    return out


def expand(volume):
    """dr_expand_cells on the host: [D,H,W] -> [D*H*W, 8] cell-major records."""
    vol = np.ascontiguousarray(volume, np.float32).reshape(np.asarray(volume).shape[-3:])
    d = make_desc(vol.shape, (8, 8), 2, 1, 0)
    ce = np.zeros(vol.size * 8, np.float32)
    lib().sim_expand(ctypes.byref(d), _p(vol), _p(ce))
    return ce.reshape(-1, 8)
