"""With a -DDR_BOUNDS_CHECK build of the library (DIFFRENDER_LIB=...), every volume load and gradient reduction of the
marches is range-checked on the device; this test drives ragged, minimal, clamped-edge and both-layout cases through it and
requires zero violations.  (compute-sanitizer is closed on the B200 pool.)  Skipped for a normal build."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def test_no_out_of_range_access_in_debug_build():
    from differender_b200 import _lib
    if "dbg" not in os.path.basename(_lib.LIB_PATH):
        # the library is chosen at import time: re-run this very test in a child process bound to the debug build
        import subprocess, sys
        from differender_b200.build import DEBUG_LIB_PATH
        if not os.path.exists(DEBUG_LIB_PATH):
            pytest.skip("libdiffrender_dbg.so not built (__graft_entry__.build() builds it)")
        env = dict(os.environ, DIFFRENDER_LIB=DEBUG_LIB_PATH)
        r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", os.path.abspath(__file__)], env=env,
                           capture_output=True, text=True, timeout=600)
        assert r.returncode == 0 and "1 passed" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
        return
    from differender_b200 import VolumeRaycaster
    from differender_b200.synthetic import make_jitter, make_tf
    dev = "cuda:0"
    lib = _lib.load()
    assert lib.dr_debug_oob_count() == 0
    g = torch.Generator().manual_seed(0)
    cams = torch.tensor([[1.2, 0.7, 2.2], [0.3, 0.2, 0.4], [-2.0, 1.5, 0.1], [0.0, 0.7, 2.5]], device=dev)   # incl. a camera inside the box
    for layout in ("linear", "brick8", "cell8"):
        for (D, H, W) in ((2, 2, 2), (5, 9, 3), (16, 16, 16), (21, 18, 27)):
            for dtype in (torch.float32, torch.float16) + ((torch.uint8,) if layout == "cell8" else ()):
                vol = torch.rand((1, D, H, W), generator=g).to(dev)
                vol[0, 0] = 1.0; vol[0, -1] = 0.0                                  # extreme intensities: TF index clamps
                vol = (vol * 255.0).to(torch.uint8) if dtype == torch.uint8 else vol.to(dtype)
                tf = make_tf("tf1", 32, device=dev).t().contiguous()[None]
                vr = VolumeRaycaster((W, D, H), (19, 13), max_samples=256, tf_resolution=32, layout=layout)
                v = vr.brick(vol.reshape(1, D, H, W).contiguous())
                jit = make_jitter(4, 13, 19, device=dev)
                for sr in (0.3, 1.0, 3.0):
                    out, K, Tp = vr.march(v, tf, cams, sr, jit)
                    vr.march_backward(v, tf, cams, sr, jit, torch.randn_like(out), out, K, Tp, True, True)
                    vr.march(v, tf, cams, sr, None, nondiff=True)
    assert lib.dr_debug_oob_count() == 0
