"""Times dr_gather_grad / dr_gather_step / dr_expand_cells alone (CUDA events, L2 flushed between launches) for the library selected by
DIFFRENDER_LIB: GB/s of compulsory traffic (32 B/voxel read + 4 B/voxel written for the gather) against the measured HBM peak."""
import ctypes, json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from differender_b200 import VolumeRaycaster, _lib
tag = sys.argv[1] if len(sys.argv) > 1 else "base"
peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6552.6
flush = torch.empty(64 << 20, dtype=torch.float32, device="cuda")
for n in (256, 512):
    vr = VolumeRaycaster((n, n, n), (64, 64), tf_resolution=16, layout="cell8")
    cells = torch.randn((1, n ** 3 * 8), device="cuda") * (torch.rand((1, n ** 3 * 8), device="cuda") < 0.3)
    out = torch.empty((1, n, n, n), device="cuda")
    p, m = torch.rand((n, n, n), device="cuda"), torch.zeros((n, n, n), device="cuda")
    vc = torch.empty((n ** 3, 8), device="cuda")
    d = vr.desc(1, 1, 1, _lib.VOX_F32, 0, 1.0)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    lib = _lib.load()

    def t(fn, reps=10):
        fn(); best = 1e9
        for _ in range(reps):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b))
        return best
    g = t(lambda: vr.gather(cells, out=out))
    gs = t(lambda: _lib.check(lib.dr_gather_step(ctypes.byref(d), _lib.ptr(cells), None, _lib.ptr(p), _lib.ptr(m), _lib.ptr(vc), None, 0.1, 0.9, 0.1, 0.0, 1.0, st), "gs"))
    gsn = t(lambda: _lib.check(lib.dr_gather_step(ctypes.byref(d), _lib.ptr(cells), None, _lib.ptr(p), _lib.ptr(m), None, None, 0.1, 0.9, 0.1, 0.0, 1.0, st), "gs"))
    ex = t(lambda: _lib.check(lib.dr_expand_cells(ctypes.byref(d), _lib.ptr(p), _lib.ptr(vc), st), "ex"))
    ms = t(lambda: _lib.check(lib.dr_momentum_step(_lib.ptr(p), _lib.ptr(out), _lib.ptr(m), n ** 3, 0.1, 0.9, 0.1, 0.0, 1.0, st), "ms"))
    v = n ** 3
    print(f"{tag} {n}^3: gather {g * 1e3:7.1f} us = {36 * v / g / 1e6:6.0f} GB/s ({100 * 36 * v / g / 1e6 / peak:4.1f} % of HBM) | gather_step {gs * 1e3:7.1f} us = {(32 + 16 + 32) * v / gs / 1e6:6.0f} GB/s "
          f"({100 * 80 * v / gs / 1e6 / peak:4.1f} %) | gather_step without the volume refresh {gsn * 1e3:7.1f} us ({100 * 48 * v / gsn / 1e6 / peak:4.1f} %) | separate: gather + momentum_step {ms * 1e3:6.1f} us + expand_cells {ex * 1e3:6.1f} us = {(g + ms + ex) * 1e3:7.1f} us", flush=True)
