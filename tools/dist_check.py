"""2-GPU check of DistributedRaycaster with NCCL: the all-reduced gradients on every rank must equal the single-process
gradients of the whole view batch.  Run:  torchrun --nproc-per-node 2 tools/dist_check.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from differender_b200 import Raycaster
from differender_b200.distributed import DistributedRaycaster, shard_views
from differender_b200.synthetic import make_cameras, make_jitter, make_tf, make_volume

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
N, w, h, R, views = 48, 64, 48, 64, 5
vol = make_volume(N).to(dev); tf = make_tf("tf1", R).to(dev)
cams = make_cameras(views).to(dev); jit = make_jitter(views, h, w).to(dev)
go = torch.randn((views, 4, h, w), generator=torch.Generator().manual_seed(0)).to(dev)
rc = Raycaster((N, N, N), (w, h), R, max_samples=1024)
v = vol.clone().requires_grad_(True); t = tf.clone().requires_grad_(True)
img, idx = DistributedRaycaster(rc)(v, t, cams, jit)
assert idx == shard_views(views, rank, world)
(img * go[idx]).sum().backward()
v2 = vol.clone().requires_grad_(True); t2 = tf.clone().requires_grad_(True)
(rc(v2, t2, cams, jit) * go).sum().backward()
ev = float((v.grad - v2.grad).norm() / v2.grad.norm()); et = float((t.grad - t2.grad).norm() / t2.grad.norm())
print(f"rank {rank}: views {idx}, rel-L2 vs single process: volume {ev:.2e}, tf {et:.2e}", flush=True)
assert ev < 1e-5 and et < 1e-5
# the same batch through FusedVolumeSGD: the gradient stays cell-major, is gathered into a flat buffer, all-reduced and applied by
# dr_gather_step from the reduced linear gradient; every rank must end up with the volume a single process computes
from differender_b200 import FusedVolumeSGD, MomentumSGD
rc_f = Raycaster((N, N, N), (w, h), R, max_samples=1024)
vf = vol.clone().requires_grad_(True)
opt = FusedVolumeSGD(rc_f, vf, lr=2.0, momentum=0.5, max_grad=0.05, hi=1.0)
imgf, _ = DistributedRaycaster(rc_f)(vf, tf, cams, jit)
(imgf * go[idx]).sum().backward()
assert vf.grad is None
opt.step()
vs = vol.clone()
MomentumSGD(vs, lr=2.0, momentum=0.5, max_grad=0.05, lo=0.0, hi=1.0).step(v2.grad)
ef = float((vf.detach() - vs).norm() / vs.norm())
print(f"rank {rank}: fused distributed volume step vs single process: rel-L2 {ef:.2e}", flush=True)
assert ef < 1e-5
dist.barrier()
dist.destroy_process_group()
