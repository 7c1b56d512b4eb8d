"""The C-ABI library loads on a CPU-only box and exports every symbol include/diffrender.h declares; host-only entry points
(descriptor construction, sizes, argument validation) work without a GPU.  No compute calls here."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "diffrender.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dr_[a-z_0-9]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    from differender_b200.build import build_library
    build_library()
    from differender_b200 import _lib
    return _lib.load()


def test_exports_every_declared_symbol(lib):
    from differender_b200 import _lib
    names = _declared_functions()
    assert set(names) == set(_lib.EXPORTS), (names, _lib.EXPORTS)
    for n in names:
        assert getattr(lib, n) is not None


def test_version_and_desc_init(lib):
    from differender_b200 import _lib
    assert lib.dr_version() == _lib.DR_VERSION
    d = _lib.make_desc(256, 256, 256, 1024, 1024, 128, 2048, 16, 1, 1, _lib.VOX_F32, _lib.F_HAS_JITTER, 1.0, 30.0, 0.1)
    assert (d.nbx, d.nby, d.nbz) == (32, 32, 32) and d.tap_generic == 0
    assert lib.dr_bricked_elems(ctypes.byref(d)) == 256 ** 3
    assert abs(d.near_h - 2 * 0.1 * 3 ** -0.5) < 1e-7              # 2*tan(30 deg)*near  (reference :146)
    assert abs(d.vol_diag - 255 * 3 ** 0.5) < 1e-3 and d.tf_len == 127.0
    assert abs(d.scale[0] - (255 - 1e-4)) < 1e-4
    assert lib.dr_workspace_bytes(ctypes.byref(d)) == 0            # no TF gradient requested
    d2 = _lib.make_desc(37, 29, 45, 50, 34, 33, 512, 2, 1, 2, _lib.VOX_F16, _lib.F_NEEDS_TF_GRAD, 0.5, 30.0, 0.1)
    assert (d2.nbx, d2.nby, d2.nbz) == (5, 4, 6) and abs(d2.inv_sr - 2.0) < 1e-7
    assert lib.dr_workspace_bytes(ctypes.byref(d2)) == 2 * 1024 * (33 + 1) * 16        # R bins + one pad bin per privatised copy
    big = _lib.make_desc(2304, 64, 64, 64, 64, 16, 64, 1, 1, 1, _lib.VOX_F32, 0, 1.0, 30.0, 0.1)
    assert big.tap_generic == 1                                     # a normal tap can skip a whole cell: generic taps


@pytest.mark.parametrize("kw,msg", [
    (dict(X=1), "volume dims"), (dict(R=1), "tf resolution"), (dict(M=0), "max_samples"), (dict(Bvol=3), "Bvol"),
    (dict(vox=7), "dtype"), (dict(sr=0.0), "sampling_rate"), (dict(X=2048, Y=2048, Z=2048), "too large"),
])
def test_desc_init_rejects_bad_arguments(lib, kw, msg):
    from differender_b200 import _lib
    a = dict(X=32, Y=32, Z=32, W=16, H=16, R=16, M=64, BS=2, Bvol=1, Btf=1, vox=0, flags=0, sr=1.0)
    a.update(kw)
    with pytest.raises(RuntimeError, match=msg):
        _lib.make_desc(a["X"], a["Y"], a["Z"], a["W"], a["H"], a["R"], a["M"], a["BS"], a["Bvol"], a["Btf"], a["vox"], a["flags"],
                       a["sr"], 30.0, 0.1)


def test_entry_points_validate_pointers_without_touching_the_gpu(lib):
    from differender_b200 import _lib
    d = _lib.make_desc(32, 32, 32, 16, 16, 16, 64, 1, 1, 1, 0, 0, 1.0, 30.0, 0.1)
    assert lib.dr_forward(ctypes.byref(d), None, None, None, None, None, None, None, None) == -1
    assert b"null pointer" in lib.dr_last_error()
    assert lib.dr_brick_volume(ctypes.byref(d), None, None, None) == -1
    assert lib.dr_expand_cells(ctypes.byref(d), None, None, None) == -1 and b"dr_expand_cells" in lib.dr_last_error()
    assert lib.dr_expand_cells(ctypes.byref(d), ctypes.c_void_p(0x1000), ctypes.c_void_p(0x1010), None) == -3      # 32-byte alignment
    assert lib.dr_gather_grad(ctypes.byref(d), None, None, 0, None) == -1
    assert lib.dr_grad_cells_elems(ctypes.byref(d)) == 32 ** 3 * 8
    d.flags = _lib.F_NEEDS_TF_GRAD
    fake = ctypes.c_void_p(0x1000)
    rc = lib.dr_backward(ctypes.byref(d), fake, fake, fake, None, fake, fake, fake, fake, None, fake, None, 0, None)
    assert rc == -5 and b"workspace" in lib.dr_last_error()
    d.flags = _lib.F_NONDIFF
    assert lib.dr_backward(ctypes.byref(d), fake, fake, fake, None, fake, fake, fake, fake, None, None, None, 0, None) == -1
    two = _lib.make_desc(32, 32, 32, 16, 16, 16, 64, 1, 1, 1, 0, _lib.F_LAYOUT_BRICK8 | _lib.F_LAYOUT_CELL8, 1.0, 30.0, 0.1)
    assert lib.dr_forward(ctypes.byref(two), fake, fake, fake, None, fake, None, None, None) == -1
    assert b"two volume layouts" in lib.dr_last_error()
    gen = _lib.make_desc(2304, 8, 8, 16, 16, 16, 64, 1, 1, 1, 0, _lib.F_LAYOUT_CELL8, 1.0, 30.0, 0.1)
    assert lib.dr_forward(ctypes.byref(gen), fake, fake, fake, None, fake, None, None, None) == -1
    assert b"linear layout" in lib.dr_last_error()                 # the generic tap path exists for the linear layout only
    zero = _lib.DrDesc()
    assert lib.dr_forward(ctypes.byref(zero), fake, fake, fake, None, fake, None, None, None) == -1
    assert b"dr_desc_init" in lib.dr_last_error()
