"""Debug aid: forward of one of the long-axis parity cases with the library selected by DIFFRENDER_LIB, saved for comparison."""
import sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from helpers import case_inputs
from differender_b200 import VolumeRaycaster
tag = sys.argv[1]
for shape in [(1100, 6, 6), (6, 1100, 6), (6, 6, 1100)]:
    for layout in ("linear", "cell8"):
        for dt in (torch.float32, torch.float16):
            vol, tf, cams, jit = case_inputs(shape, (24, 20), 32, seed=5, views=1)
            D, H, W = shape
            vr = VolumeRaycaster((W, D, H), (24, 20), max_samples=4096, tf_resolution=32, layout=layout)
            b = vr.brick(vol.cuda().to(dt).reshape(1, D, H, W).contiguous())
            out, K, Tp = vr.march(b, tf.cuda().t().contiguous()[None], cams.cuda().contiguous(), 1.0, jit.cuda().contiguous())
            np.save(f"gpurun_out/cmp_{tag}_{shape[0]}_{shape[1]}_{layout}_{str(dt)[-2:]}.npy", out.cpu().numpy())
