"""ctypes binding of libdiffrender.so (include/diffrender.h).  There is no CPU fallback: if the library is missing
or a call fails, a RuntimeError is raised."""
import ctypes
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
# DIFFRENDER_LIB selects another build of the same library (tuning experiments); default: the in-tree build
LIB_PATH = os.environ.get("DIFFRENDER_LIB") or os.path.join(_PKG, "libdiffrender.so")

# include/diffrender.h
DR_VERSION = 104
VOX_F32, VOX_F16, VOX_U8 = 0, 1, 2
F_NONDIFF, F_NEEDS_VOL_GRAD, F_NEEDS_TF_GRAD, F_HAS_JITTER, F_OUT_IMAGE, F_TF_4R, F_GENERIC_TAPS, F_LAYOUT_BRICK8, F_COUNT_SHADED, F_LAYOUT_CELL8 = 1, 2, 4, 8, 16, 32, 64, 256, 512, 2048

EXPORTS = ("dr_version", "dr_debug_oob_count", "dr_last_error", "dr_desc_init", "dr_bricked_elems", "dr_brick_volume", "dr_expand_cells", "dr_forward",
           "dr_workspace_bytes", "dr_grad_cells_elems", "dr_backward", "dr_gather_grad", "dr_forward_mse", "dr_backward_mse",
           "dr_skip_minmax_bytes", "dr_skip_grid_bytes", "dr_build_skip_grid", "dr_forward_ex",
           "dr_momentum_step", "dr_ingest_u8", "dr_gather_step", "dr_probe_l2_read", "dr_backward_ex")


class DrDesc(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in ("X", "Y", "Z", "W", "H", "R", "M", "BS", "Bvol", "Btf", "vox_dtype")] + \
               [("flags", ctypes.c_uint32)] + \
               [(n, ctypes.c_float) for n in ("sr", "inv_sr", "near_", "near_w", "near_h")] + \
               [("scale", ctypes.c_float * 3)] + \
               [(n, ctypes.c_float) for n in ("vol_diag", "tf_len", "ambient", "diffuse", "specular", "ert", "delta",
                                              "alpha_skip")] + \
               [(n, ctypes.c_int32) for n in ("nbx", "nby", "nbz", "tap_generic")]


_lib = None


def load():
    """Loads the library once.  Raises RuntimeError (never falls back) when it is missing or stale."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"differender_b200: {LIB_PATH} not found. Build it with `python -m differender_b200.build` "
            "(nvcc, sm_100a). There is no CPU or PyTorch fallback for the ray-march.")
    lib = ctypes.CDLL(LIB_PATH)
    vp, cp = ctypes.c_void_p, ctypes.c_char_p
    dp = ctypes.POINTER(DrDesc)
    lib.dr_version.restype = ctypes.c_int
    lib.dr_debug_oob_count.restype = ctypes.c_longlong
    lib.dr_last_error.restype = cp
    lib.dr_desc_init.argtypes = [dp] + [ctypes.c_int32] * 11 + [ctypes.c_uint32] + [ctypes.c_double] * 3
    lib.dr_desc_init.restype = ctypes.c_int
    lib.dr_bricked_elems.argtypes = [dp]; lib.dr_bricked_elems.restype = ctypes.c_size_t
    lib.dr_workspace_bytes.argtypes = [dp]; lib.dr_workspace_bytes.restype = ctypes.c_size_t
    lib.dr_brick_volume.argtypes = [dp, vp, vp, vp]; lib.dr_brick_volume.restype = ctypes.c_int
    lib.dr_expand_cells.argtypes = [dp, vp, vp, vp]; lib.dr_expand_cells.restype = ctypes.c_int
    lib.dr_forward.argtypes = [dp, vp, vp, vp, vp, vp, vp, vp, vp]; lib.dr_forward.restype = ctypes.c_int
    lib.dr_backward.argtypes = [dp] + [vp] * 11 + [ctypes.c_size_t, vp]; lib.dr_backward.restype = ctypes.c_int
    lib.dr_grad_cells_elems.argtypes = [dp]; lib.dr_grad_cells_elems.restype = ctypes.c_size_t
    lib.dr_gather_grad.argtypes = [dp, vp, vp, ctypes.c_int, vp]; lib.dr_gather_grad.restype = ctypes.c_int
    lib.dr_forward_mse.argtypes = [dp] + [vp] * 10; lib.dr_forward_mse.restype = ctypes.c_int
    lib.dr_forward_ex.argtypes = [dp] + [vp] * 11; lib.dr_forward_ex.restype = ctypes.c_int
    lib.dr_skip_minmax_bytes.argtypes = [dp]; lib.dr_skip_minmax_bytes.restype = ctypes.c_size_t
    lib.dr_skip_grid_bytes.argtypes = [dp]; lib.dr_skip_grid_bytes.restype = ctypes.c_size_t
    lib.dr_build_skip_grid.argtypes = [dp, vp, vp, vp, ctypes.c_int, vp, vp]; lib.dr_build_skip_grid.restype = ctypes.c_int
    lib.dr_backward_mse.argtypes = [dp] + [vp] * 5 + [ctypes.c_float] + [vp] * 6 + [ctypes.c_size_t, vp]
    lib.dr_backward_mse.restype = ctypes.c_int
    lib.dr_backward_ex.argtypes = [dp] + [vp] * 6 + [ctypes.c_float] + [vp] * 8 + [ctypes.c_size_t, vp]
    lib.dr_backward_ex.restype = ctypes.c_int
    lib.dr_momentum_step.argtypes = [vp, vp, vp, ctypes.c_size_t] + [ctypes.c_float] * 5 + [vp]
    lib.dr_momentum_step.restype = ctypes.c_int
    lib.dr_ingest_u8.argtypes = [dp, vp, vp, ctypes.c_int, vp]; lib.dr_ingest_u8.restype = ctypes.c_int
    lib.dr_gather_step.argtypes = [dp] + [vp] * 6 + [ctypes.c_float] * 5 + [vp]; lib.dr_gather_step.restype = ctypes.c_int
    lib.dr_probe_l2_read.argtypes = [vp, ctypes.c_size_t, ctypes.c_int, vp, vp]; lib.dr_probe_l2_read.restype = ctypes.c_longlong
    if lib.dr_version() != DR_VERSION:
        raise RuntimeError(f"differender_b200: libdiffrender.so version {lib.dr_version()} != binding {DR_VERSION}; rebuild")
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().dr_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"differender_b200: {what} failed ({rc}): {msg}")


def make_desc(X, Y, Z, W, H, R, M, BS, Bvol, Btf, vox_dtype, flags, sampling_rate, fov, near):
    d = DrDesc()
    check(load().dr_desc_init(ctypes.byref(d), X, Y, Z, W, H, R, M, BS, Bvol, Btf, vox_dtype, flags,
                              float(sampling_rate), float(fov), float(near)), "dr_desc_init")
    return d


def ptr(t):
    """Device pointer of a tensor-like (anything with data_ptr()) or None."""
    return None if t is None else ctypes.c_void_p(t.data_ptr())
