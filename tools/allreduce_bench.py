"""All-reduce of the gradient buffer sizes of C3 / C4 / C5 (64 MiB, 512 MiB, 4 GiB fp32) over the GPUs of one box: NCCL as
configured by the environment, and -- when this torch build offers it -- the symmetric-memory (NVLS multimem / two-shot) all-reduce.
    torchrun --nproc-per-node N tools/allreduce_bench.py [tag]
Prints one line per (method, size): time (barrier first, CUDA events, max over ranks), algorithm and bus bandwidth vs 900 GB/s."""
import os, sys, json
import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
tag = sys.argv[1] if len(sys.argv) > 1 else "default"
sizes = [64 << 20, 512 << 20, 4 << 30]
out = []


def timed(fn, reps):
    for _ in range(2):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    ms = torch.tensor([a.elapsed_time(b) / reps], dtype=torch.float64, device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms[0])


for nbytes in sizes:
    buf = torch.ones(nbytes // 4, dtype=torch.float32, device=dev)
    ms = timed(lambda: dist.all_reduce(buf), 5 if nbytes < (1 << 30) else 3)
    out.append(("nccl:" + tag, nbytes, ms))
    del buf
try:
    import torch.distributed._symmetric_memory as symm
    for nbytes in sizes[:2]:
        t = symm.empty(nbytes // 4, dtype=torch.float32, device=dev)
        symm.rendezvous(t, dist.group.WORLD.group_name)
        t.fill_(1.0)
        for name in ("multimem_all_reduce_", "two_shot_all_reduce_"):
            op = getattr(torch.ops.symm_mem, name, None)
            if op is None:
                continue
            try:
                ms = timed(lambda: op(t, "sum", dist.group.WORLD.group_name), 5)
                out.append(("symm_mem." + name, nbytes, ms))
            except Exception as e:
                if rank == 0:
                    print(f"symm_mem.{name} at {nbytes >> 20} MiB failed: {type(e).__name__}: {str(e)[:200]}", flush=True)
        del t
except Exception as e:
    if rank == 0:
        print(f"symmetric memory unavailable: {type(e).__name__}: {str(e)[:200]}", flush=True)
if rank == 0:
    for m, nbytes, ms in out:
        bus = nbytes * 2 * (world - 1) / world / (ms * 1e-3) / 1e9
        print(f"{m:32s} N={world} {nbytes >> 20:5d} MiB  {ms:8.3f} ms  alg {nbytes / (ms * 1e-3) / 1e9:7.1f} GB/s  bus {bus:7.1f} GB/s  ({100 * bus / 900:4.1f} % of 900)", flush=True)
dist.barrier()
dist.destroy_process_group()
