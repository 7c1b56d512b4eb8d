"""Diagnostic: per-step wall time of the autograd e2e loop next to the caching allocator's cudaMalloc/cudaFree counts."""
import gc, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from differender_b200 import Raycaster
from differender_b200.synthetic import make_cameras, make_jitter, make_tf, make_volume
dev = torch.device("cuda", 0)
n, w, h, R, views = 256, 1024, 1024, 128, 16
vol = make_volume(n, device=dev); tf = make_tf("tf1", R, device=dev); cams = make_cameras(views, device=dev)
jit = make_jitter(views, h, w, device=dev); target = torch.rand((views, 4, h, w), device=dev)
rc = Raycaster((n, n, n), (w, h), R, sampling_rate=1.0, jitter=True, max_samples=2048, layout=sys.argv[1] if len(sys.argv) > 1 else "auto")
mode = sys.argv[2] if len(sys.argv) > 2 else "plain"
if mode == "nogc":
    gc.disable()
for i in range(24):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    s0 = torch.cuda.memory_stats()
    v = vol.clone().requires_grad_(True); t = tf.clone().requires_grad_(True)
    img = rc(v, t, cams, jit)
    loss = ((img - target) ** 2).mean()
    loss.backward()
    l = float(loss)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    s1 = torch.cuda.memory_stats()
    del v, t, img, loss
    if mode == "collect":
        gc.collect()
    print(f"step {i:2d} {1e3 * (t1 - t0):7.2f} ms  cudaMalloc +{s1['num_device_alloc'] - s0['num_device_alloc']} cudaFree +{s1['num_device_free'] - s0['num_device_free']} "
          f"reserved {s1['reserved_bytes.all.current'] / 2**30:.2f} GiB  retries {s1['num_alloc_retries']}")
