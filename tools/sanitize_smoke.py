"""Small end-to-end run for compute-sanitizer (memcheck / racecheck): forward + backward, both layouts, fp32 and fp16, nondiff,
ragged sizes.  Run:  compute-sanitizer --tool memcheck python tools/sanitize_smoke.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from differender_b200 import VolumeRaycaster, MomentumSGD
from differender_b200.synthetic import make_cameras, make_jitter, make_tf, make_volume
from differender_b200.utils import volume_from_raw_u8

dev = "cuda:0"
for layout in ("linear", "brick8"):
    for dtype in (torch.float32, torch.float16):
        D, H, W = 21, 18, 27
        vol = make_volume((D, H, W), device=dev, dtype=dtype)
        tf = make_tf("tf1", 64, device=dev).t().contiguous()[None]
        cams = make_cameras(2, device=dev)
        jit = make_jitter(2, 22, 30, device=dev)
        vr = VolumeRaycaster((W, D, H), (30, 22), max_samples=512, tf_resolution=64, layout=layout)
        v = vr.brick(vol.reshape(1, D, H, W))
        out, K, Tp = vr.march(v, tf, cams, 1.0, jit)
        target = torch.rand_like(out)
        out2, K2, Tp2, ls = vr.march(v, tf, cams, 1.0, jit, mse_target=target)
        go = torch.randn_like(out)
        gv, gt = vr.march_backward(v, tf, cams, 1.0, jit, go, out, K, Tp, True, True)
        gv2, gt2 = vr.march_backward(v, tf, cams, 1.0, jit, target, out, K, Tp, True, True, mse_scale=0.01)
        nd, _, _ = vr.march(v, tf, cams, 4.0, None, nondiff=True)
        torch.cuda.synchronize()
        print(layout, dtype, float(out.sum()), float(gv.abs().sum()), float(gt.abs().sum()), float(ls.sum()), float(nd.sum()))
p = torch.rand(4, 64, device=dev)
MomentumSGD(p).step(torch.randn(4, 64, device=dev))
raw = torch.randint(0, 256, (9, 7, 11), dtype=torch.uint8)
print(volume_from_raw_u8(raw).shape)
torch.cuda.synchronize()
print("sanitize smoke done")
