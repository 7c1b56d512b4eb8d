"""Builds libdiffrender.so (the sm_100a kernels + C ABI) in-tree with nvcc.  No JIT, no torch extension machinery:
the library has no torch types in its interface (include/diffrender.h).

The march kernels are instantiated in separate translation units (csrc/dr_fwd_f32.cu, dr_fwd_f16.cu, dr_bwd_f32.cu,
dr_bwd_f16.cu) that are
compiled in parallel and linked with csrc/diffrender.cu (C ABI + the small kernels)."""
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
_CSRC = os.path.join(_PKG, "csrc")
LIB_PATH = os.path.join(_PKG, "libdiffrender.so")
DEBUG_LIB_PATH = os.path.join(_PKG, "libdiffrender_dbg.so")      # -DDR_BOUNDS_CHECK build used by tests/test_gpu_bounds.py
UNITS = ["diffrender.cu", "dr_fwd_f32.cu", "dr_fwd_f16.cu", "dr_fwd_u8.cu", "dr_bwd_f32.cu", "dr_bwd_f16.cu", "dr_bwd_u8.cu"]
SOURCES = [os.path.join(_CSRC, u) for u in UNITS]
HEADERS = [os.path.join(_CSRC, h) for h in ("dr_math.cuh", "dr_kernels.cuh", "dr_host.h", "dr_desc.h")] + \
          [os.path.join(_ROOT, "include", "diffrender.h")]
NVCC_FLAGS = ["-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3",
              "-Xcompiler", "-fPIC", "-ccbin", "/usr/bin/g++"]


def needs_build(path=LIB_PATH):
    if not os.path.exists(path):
        return True
    t = os.path.getmtime(path)
    return any(os.path.getmtime(p) > t for p in SOURCES + HEADERS)


def build_library(force=False, verbose=False, debug=False, out=None, defines=()):
    """Compile for sm_100a only (nvcc cross-compiles without a GPU).  Returns the path of the .so.
    debug=True builds the bounds-checking variant (every volume load / gradient reduction range-checked on the device) as
    ONE translation unit (its violation counter is a single __device__ variable).  `out` / `defines` build a tuning
    variant (e.g. defines=("DR_BWD_MIN_BLOCKS=5",)) that bench.py can select with DIFFRENDER_LIB."""
    out = out or (DEBUG_LIB_PATH if debug else LIB_PATH)
    if not force and not needs_build(out):
        return out
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    common = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-D" + d for d in defines] + \
             ["-I" + os.path.join(_ROOT, "include"), "-I" + _CSRC]
    if debug:
        subprocess.check_call(common + ["-DDR_BOUNDS_CHECK", "-DDR_UNITY_BUILD", "--shared", "-o", out, SOURCES[0]])
        return out
    objdir = os.path.join(_PKG, "build", os.path.basename(out) + ".o")
    os.makedirs(objdir, exist_ok=True)
    objs = [os.path.join(objdir, u.replace(".cu", ".o")) for u in UNITS]

    def compile_unit(i):
        r = subprocess.run(common + ["-c", "-o", objs[i], SOURCES[i]], capture_output=True, text=True)
        return r.returncode, r.stdout + r.stderr

    with ThreadPoolExecutor(len(UNITS)) as ex:
        results = list(ex.map(compile_unit, range(len(UNITS))))
    for (rc, log), u in zip(results, UNITS):
        if verbose or rc:
            print(f"---- {u}\n{log}")
        if rc:
            raise subprocess.CalledProcessError(rc, f"nvcc -c {u}")
    subprocess.check_call([nvcc, "--shared", "-gencode", "arch=compute_100a,code=sm_100a", "-ccbin", "/usr/bin/g++", "-o", out] + objs)
    return out


if __name__ == "__main__":
    print(build_library(force=True, verbose=True))
