"""Builds libdiffrender.so (the sm_100a kernels + C ABI) in-tree with nvcc.  No JIT, no torch extension machinery:
the library has no torch types in its interface (include/diffrender.h)."""
import os
import subprocess

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
LIB_PATH = os.path.join(_PKG, "libdiffrender.so")
DEBUG_LIB_PATH = os.path.join(_PKG, "libdiffrender_dbg.so")      # -DDR_BOUNDS_CHECK build used by tests/test_gpu_bounds.py
SOURCES = [os.path.join(_PKG, "csrc", "diffrender.cu")]
HEADERS = [os.path.join(_PKG, "csrc", "dr_math.cuh"), os.path.join(_PKG, "csrc", "dr_desc.h"),
           os.path.join(_ROOT, "include", "diffrender.h")]
NVCC_FLAGS = ["-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3",
              "--shared", "-Xcompiler", "-fPIC", "-ccbin", "/usr/bin/g++"]


def needs_build(path=LIB_PATH):
    if not os.path.exists(path):
        return True
    t = os.path.getmtime(path)
    return any(os.path.getmtime(p) > t for p in SOURCES + HEADERS)


def build_library(force=False, verbose=False, debug=False):
    """Compile for sm_100a only (nvcc cross-compiles without a GPU).  Returns the path of the .so.
    debug=True builds the bounds-checking variant (every volume load / gradient reduction range-checked on the device)."""
    out = DEBUG_LIB_PATH if debug else LIB_PATH
    if not force and not needs_build(out):
        return out
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + (["-DDR_BOUNDS_CHECK"] if debug else []) + \
          ["-I" + os.path.join(_ROOT, "include"), "-I" + os.path.join(_PKG, "csrc"), "-o", out] + SOURCES
    subprocess.check_call(cmd)
    return out


if __name__ == "__main__":
    print(build_library(force=True, verbose=True))
