import os, sys, torch
sys.path.insert(0, os.getcwd())
from differender_b200 import _lib, VolumeRaycaster
from differender_b200.synthetic import make_jitter, make_tf
dev = "cuda:0"
lib = _lib.load()
g = torch.Generator().manual_seed(0)
cams = torch.tensor([[1.2, 0.7, 2.2], [0.3, 0.2, 0.4], [-2.0, 1.5, 0.1], [0.0, 0.7, 2.5]], device=dev)
last = lib.dr_debug_oob_count()
def chk(tag):
    global last
    c = lib.dr_debug_oob_count()
    if c != last: print(tag, "+", c - last)
    last = c
for skip in (False, True):
  for layout in ("linear", "brick8", "cell8"):
    for (D, H, W) in ((2, 2, 2), (5, 9, 3), (16, 16, 16), (21, 18, 27)):
        for dtype in (torch.float32, torch.float16):
            vol = torch.rand((1, D, H, W), generator=g).to(dev, dtype)
            vol[0, 0] = 1.0; vol[0, -1] = 0.0
            tf = make_tf("tf1", 32, device=dev).t().contiguous()[None]
            vr = VolumeRaycaster((W, D, H), (19, 13), max_samples=256, tf_resolution=32, layout=layout, skip_empty=skip)
            v = vr.brick(vol.reshape(1, D, H, W).contiguous())
            jit = make_jitter(4, 13, 19, device=dev)
            for sr in (0.3, 1.0, 3.0):
                out, K, Tp = vr.march(v, tf, cams, sr, jit); chk(f"skip={skip} {layout} {(D,H,W)} {dtype} sr={sr} fwd")
                vr.march_backward(v, tf, cams, sr, jit, torch.randn_like(out), out, K, Tp, True, True); chk(f"skip={skip} {layout} {(D,H,W)} {dtype} sr={sr} bwd")
                vr.march(v, tf, cams, sr, None, nondiff=True); chk(f"skip={skip} {layout} {(D,H,W)} {dtype} sr={sr} nondiff")
print("total", lib.dr_debug_oob_count())
