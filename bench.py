#!/usr/bin/env python
"""Benchmark of the differentiable ray-march hot path (BASELINE.json metric: Gsamples/s forward and forward+backward).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config c3] [--impl ours|reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

One "step" = one pass of the hot path over one batch of views: cell-major copy of the volume, skip grid (per-macro-cell
min/max + TF classification), forward march, backward march (TF + volume gradients), gather of the cell-major gradient (and,
for N > 1, all-reduce [volume grad | TF grad]).  A "sample" is one ACTIVE ray-march step; the count is the sum
of the forward kernel's per-ray K -- identical with and without the exact empty-space skipping (--no-skip).

Prints ONE JSON line on rank 0.  `value` = device-resident throughput through the C ABI; `e2e` = the same metric through
the public `Raycaster` autograd API with all inputs copied from pinned host memory every step (double-buffered: step i+1's
copies run on a second stream under step i's compute).
`--impl reference` times the CPU oracle (oracle/cpu_ref.c, kind "port": the real reference needs Taichi, which is not
installable here) on the host cores on a bounded sample of the same workload.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # name: volume N^3, dtype, image (w,h), views per GPU, mode
    "c1": dict(n=256, dtype="f32", res=(512, 512), views=1, mode="nondiff", sr=16.0, M=1, jitter=False,
               desc="C1 forward-only nondiff render, 256^3 fp32, 512x512, 1 view, sr 16"),
    "c2": dict(n=256, dtype="f32", res=(512, 512), views=1, mode="tf", sr=1.0, M=2048, jitter=True, tf="black",
               desc="C2 TF optimisation step: fwd+bwd w.r.t. TF only + momentum update, 256^3 fp32, 512x512, 1 view"),
    "c3": dict(n=256, dtype="f32", res=(1024, 1024), views=16, mode="full", sr=1.0, M=2048, jitter=True,
               desc="C3 volume-gradient backprop: fwd+bwd (TF+volume grad), 256^3 fp32, 1024x1024, 16 views per GPU"),
    "c4": dict(n=512, dtype="f32", res=(1024, 1024), views=8, mode="full", sr=1.0, M=4096, jitter=True,
               desc="C4 multi-view batch: fwd+bwd, 512^3 fp32, 1024x1024, 8 views per GPU (64 over 8 GPUs)"),
    "c5": dict(n=1024, dtype="f16", res=(2048, 2048), views=32, mode="full", sr=1.0, M=8192, jitter=True,
               desc="C5 large volume: fwd+bwd, 1024^3 fp16-stored, 2048x2048, jittered, 32 views per GPU (256 over 8 GPUs)"),
}
L2_FLUSH_BYTES = 256 << 20
METRIC = {"full": "Gsamples/s fwd+bwd (TF+volume grad)", "tf": "Gsamples/s fwd+bwd (TF grad)", "nondiff": "Gsamples/s fwd"}


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=5)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--config", default="c3", choices=sorted(CONFIGS))
    p.add_argument("--views", type=int, default=None, help="views per GPU (default: the config's)")
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--tf", default=None, help="transfer-function preset (reference utils.get_tf): tf1..tf5, gray, black, rand; "
                                              "default tf1 (C2: its optimisation start `black`)")
    p.add_argument("--layout", default="auto", choices=["auto", "linear", "brick8", "cell8"], help="volume layout read by the march kernels")
    p.add_argument("--no-skip", action="store_true", help="march every sample (no exact empty-space skip grid in the forward)")
    p.add_argument("--cuda-profiler-range", action="store_true",
                   help="wrap the timed region in cudaProfilerStart/Stop (for ncu --profile-from-start off)")
    return p.parse_args()


# ------------------------------------------------------------------------------------------------------------------
# clocks (nvidia-smi sampled DURING the timed region)
# ------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for (t, r) in self.rows if t0 <= t <= t1 + 0.2] or [r for (_, r) in self.rows]
        sm, mx, reasons = [], None, set()
        for r in rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------------------
# CPU oracle leg (cpu_baseline / --impl reference)
# ------------------------------------------------------------------------------------------------------------------
def cpu_sample(cfg, budget_s=20.0):
    """Times the CPU restatement on a bounded sample of the workload: view 0 of the batch at the workload's ray count
    (volume capped at 256^3).  Returns dict(value Gsamples/s, seconds, samples, cores, sample description)."""
    import numpy as np
    import torch
    from differender_b200.synthetic import make_cameras, make_jitter, make_tf, make_volume
    from oracle import cpu_oracle as co
    co.build()
    n = min(cfg["n"], 256)                      # the CPU leg keeps the volume at <= 256^3 (memory/time); per-sample work is size-independent
    w, h = min(cfg["res"][0], 1024), min(cfg["res"][1], 1024)
    vol = make_volume(n).numpy()
    if cfg["dtype"] == "f16":
        vol = vol.astype(np.float16).astype(np.float32)
    tf = make_tf(cfg.get("tf", "tf1"), 128).numpy()
    cam = make_cameras(max(cfg["views"], 1))[0].numpy()
    jit = make_jitter(1, h, w)[0].numpy() if cfg["jitter"] else None
    nondiff = cfg["mode"] == "nondiff"
    kw = dict(sampling_rate=cfg["sr"], max_samples=max(cfg["M"], 1), jitter=jit, fast=True)
    co.set_num_threads(os.cpu_count() or 1, fast=True)      # torchrun exports OMP_NUM_THREADS=1; use every host core
    cores = co.num_threads(fast=True)
    t0 = time.perf_counter()
    img, K, _ = co.forward(vol, tf, cam, (w, h), nondiff=nondiff, return_counts=True, **kw)
    t_f = time.perf_counter() - t0
    samples = int(K.sum())
    t_b = 0.0
    if not nondiff:
        go = (2.0 * (img - 0.5) / img.size).astype(np.float32)
        t0 = time.perf_counter()
        co.backward(vol, tf, cam, go, (w, h), want_vol=cfg["mode"] == "full", want_tf=True, **kw)
        t_b = time.perf_counter() - t0
    return dict(value=samples / (t_f + t_b) / 1e9, fwd_value=samples / t_f / 1e9, seconds=t_f + t_b, samples=samples, cores=cores,
                sample=f"view 0 of the workload's {max(cfg['views'], 1)} views on a {n}^3 volume at {w}x{h} rays, "
                       f"{'forward' if nondiff else 'forward+backward'}, {samples} active samples, {t_f + t_b:.1f} s")


def run_reference(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals, last = [], None
    for i in range(args.warmup + args.steps):
        last = cpu_sample(cfg)
        if i >= args.warmup:
            vals.append(last)
    secs = sum(v["seconds"] for v in vals)
    samples = sum(v["samples"] for v in vals)
    value = samples / secs / 1e9
    line = {
        "impl": "reference", "metric": METRIC[cfg["mode"]],
        "value": value, "unit": "Gsamples/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * secs / max(args.steps, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": cfg["desc"], "note": "CPU restatement of the reference (oracle/cpu_ref.c); Taichi ti.cpu is not installable here"},
        "cpu_baseline": {"value": value, "unit": "Gsamples/s", "cores": last["cores"], "kind": "port", "sample": last["sample"]},
        "e2e": {"value": value, "unit": "Gsamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------------------
def run_ours(args, cfg):
    import torch
    import torch.distributed as dist
    from differender_b200 import Raycaster, VolumeRaycaster
    from differender_b200.synthetic import make_cameras, make_jitter, make_tf, make_volume

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    n, (w, h), R, M, sr = cfg["n"], cfg["res"], 128, cfg["M"], cfg["sr"]
    views = args.views or cfg["views"]
    mode = cfg["mode"]
    vdtype = torch.float16 if cfg["dtype"] == "f16" else torch.float32
    vol = make_volume(n, device=dev, dtype=vdtype)                                   # (1, D, H, W), replicated on every rank
    tf_name = args.tf or cfg.get("tf", "tf1")
    tf = make_tf(tf_name, R, device=dev)                                             # (4, R)
    all_cams = make_cameras(views * world, device=dev)
    cams = all_cams[rank * views:(rank + 1) * views].contiguous()                   # this rank's shard of the view batch
    jit = make_jitter(views, h, w, seed=4321 + rank, device=dev) if cfg["jitter"] else None
    g = torch.Generator(device=dev).manual_seed(99 + rank)
    target = torch.rand((views, 4, h, w), generator=g, device=dev)

    vr = VolumeRaycaster((n, n, n), (w, h), max_samples=M, tf_resolution=R, layout=args.layout, skip_empty=not args.no_skip)
    vol_lin = vol.reshape(1, n, n, n)
    tf_r4 = tf.t().contiguous()[None]
    flush = torch.empty(L2_FLUSH_BYTES // 4, dtype=torch.float32, device=dev)
    need_vol, need_tf = mode == "full", mode in ("full", "tf")
    momentum = torch.zeros_like(tf)
    flat_grad = torch.empty(n ** 3 + R * 4, dtype=torch.float32, device=dev) if world > 1 else None
    ev = lambda: torch.cuda.Event(enable_timing=True)
    phase_ms = {"brick": 0.0, "fwd": 0.0, "bwd": 0.0, "post": 0.0}
    samples_per_step = [0]
    launches = [0]

    def step(timed):
        e = [ev() for _ in range(5)] if timed else None
        flush.zero_()                                                                # L2 flush between steps
        if timed: e[0].record()
        if need_vol:
            vr.forget_volume()                                                       # a volume-gradient loop changes the volume every step: rebuild the skip grid's min/max too
        bricked = vr.brick(vol_lin)
        if timed: e[1].record()
        out, K, Tp = vr.march(bricked, tf_r4, cams, sr, jit, nondiff=mode == "nondiff")
        if timed: e[2].record()
        n_k = (2 if bricked.ndim in (2, 3) else 1) + (0 if args.no_skip else 2)      # (brick_kernel / expand_cells_kernel +) (skip_minmax + skip_classify +) fwd_kernel
        if mode != "nondiff":
            go = (2.0 / out.numel()) * (out - target)                                # MSE gradient
            xf = 0
            if world > 1 and need_vol:
                # the gather writes straight into the flat [volume grad | TF grad] buffer that is all-reduced (no concatenation copy)
                cells = torch.zeros((1, n ** 3 * 8), dtype=torch.float32, device=dev)
                _, gtf = vr.march_backward(bricked, tf_r4, cams, sr, jit, go, out, K, Tp, need_vol, need_tf, grad_cells=cells, extra_flags=xf)
                gvol = vr.gather(cells, out=flat_grad[:n ** 3].view(1, n, n, n))
                flat_grad[n ** 3:].copy_(gtf.reshape(-1))
            else:
                gvol, gtf = vr.march_backward(bricked, tf_r4, cams, sr, jit, go, out, K, Tp, need_vol, need_tf, extra_flags=xf)
            if timed: e[3].record()
            n_k += 1 + (1 if need_tf else 0) + (1 if need_vol else 0)                 # bwd_kernel (+ tf_reduce_kernel) (+ gather_grad_kernel)
            if world > 1:
                if not need_vol:
                    flat_grad[:gtf.numel()].copy_(gtf.reshape(-1))
                dist.all_reduce(flat_grad if need_vol else flat_grad[:gtf.numel()])      # ONE collective: NCCL over NVLink
            if mode == "tf":                                                         # C2: momentum update (reference example :375-381)
                gt = gtf[0].t().clamp(-0.1, 0.1)
                momentum.mul_(0.9).add_(gt, alpha=0.1)
        elif timed:
            e[3].record()
        if timed:
            e[4].record()
        return e, K, n_k

    for _ in range(max(args.warmup, 3)):
        _, K, n_k = step(False)
    samples_per_step[0] = int(K.sum().item())
    # diagnostic (untimed): how many of the active samples have non-zero opacity, i.e. are actually shaded
    _, Ksh, _ = vr.march(vr.brick(vol_lin), tf_r4, cams, sr, jit, nondiff=mode == "nondiff", extra_flags=512)
    shaded_fraction = float(Ksh.sum().item()) / max(samples_per_step[0], 1)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.25)
    torch.cuda.synchronize()
    t_wall0 = time.time()
    if args.cuda_profiler_range:
        torch.cuda.profiler.start()
    t0, t1 = ev(), ev()
    t0.record()
    evs = []
    for _ in range(args.steps):
        e, K, n_k = step(True)
        evs.append(e)
        launches[0] += n_k
    t1.record()
    torch.cuda.synchronize()
    if args.cuda_profiler_range:
        torch.cuda.profiler.stop()
    if world > 1:
        dist.barrier()
    t_wall1 = time.time()
    clocks = sampler.stop(t_wall0, t_wall1)
    total_ms = t0.elapsed_time(t1)
    for e in evs:
        for name, a, b in (("brick", 0, 1), ("fwd", 1, 2), ("bwd", 2, 3), ("post", 3, 4)):
            phase_ms[name] += e[a].elapsed_time(e[b])
    tot = torch.tensor([total_ms, float(samples_per_step[0])], dtype=torch.float64, device=dev)
    if world > 1:
        mx = tot.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = tot.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        total_ms, all_samples = float(mx[0]), float(sm[1])
    else:
        all_samples = float(tot[1])
    value = all_samples * args.steps / (total_ms * 1e-3) / 1e9

    # ---- end-to-end through the public autograd API, inputs from pinned host memory every step ---------------------
    e2e = None
    if not args.no_e2e:
        rc = Raycaster((n, n, n), (w, h), R, sampling_rate=sr, jitter=cfg["jitter"], max_samples=M, layout=args.layout, skip_empty=not args.no_skip)
        pin = lambda t: t.detach().cpu().pin_memory()
        h_vol, h_tf, h_cams, h_target = pin(vol), pin(tf), pin(cams), pin(target)
        h_jit = pin(jit) if jit is not None else None
        h_loss = torch.empty(1, dtype=torch.float32).pin_memory()
        h_gtf = torch.empty((4, R), dtype=torch.float32).pin_memory()
        h2d = sum(t.numel() * t.element_size() for t in (h_vol, h_tf, h_cams, h_target) + ((h_jit,) if h_jit is not None else ()))
        d2h = 4 + (h_gtf.numel() * 4 if need_tf else 0)

        e2e_phase = {"wait_for_inputs": 0.0, "forward": 0.0, "loss+backward": 0.0, "d2h+sync": 0.0}
        # Inputs are double-buffered like a data loader would: while step i computes, the copy engine brings step i+1's inputs
        # (volume, TF, cameras, jitter, target: every step copies all of them from pinned host memory) into the other buffer set
        # on a second stream.  Every timed step issues exactly one such set of copies inside the timed region.
        copy_stream = torch.cuda.Stream(device=dev)
        bufs = [dict(v=torch.empty_like(vol), t=torch.empty_like(tf), c=torch.empty_like(cams), tg=torch.empty_like(target),
                     j=torch.empty_like(jit) if jit is not None else None) for _ in range(2)]
        ready = [torch.cuda.Event(), torch.cuda.Event()]
        step_no = [0]

        def prefetch(k):
            copy_stream.wait_stream(torch.cuda.current_stream())      # the set's previous consumer (two steps back) is done
            with torch.cuda.stream(copy_stream):
                bb = bufs[k]
                bb["v"].copy_(h_vol, non_blocking=True); bb["t"].copy_(h_tf, non_blocking=True); bb["c"].copy_(h_cams, non_blocking=True)
                if bb["j"] is not None:
                    bb["j"].copy_(h_jit, non_blocking=True)
                bb["tg"].copy_(h_target, non_blocking=True)
                ready[k].record(copy_stream)

        prefetch(0)

        def e2e_step(timed=False):
            e = [ev() for _ in range(5)] if timed else None
            k = step_no[0] & 1
            step_no[0] += 1
            flush.zero_()
            if timed: e[0].record()
            main = torch.cuda.current_stream()
            main.wait_event(ready[k])                                 # this step's inputs (copied while the previous step ran)
            prefetch(k ^ 1)                                           # next step's inputs, under this step's compute
            if timed: e[1].record()
            bb = bufs[k]
            v, t, c, j, tg = bb["v"].detach(), bb["t"].detach(), bb["c"], bb["j"], bb["tg"]
            if mode == "nondiff":
                img = rc.raycast_nondiff(v, t, c, sampling_rate=sr)
                if timed: e[2].record()
                loss = ((img - tg) ** 2).mean()
            else:
                v.requires_grad_(need_vol); t.requires_grad_(need_tf)
                img = rc(v, t, c, j)
                if timed: e[2].record()
                loss = ((img - tg) ** 2).mean()
                loss.backward()
                if world > 1:
                    flat = torch.cat([x.grad.reshape(-1).float() for x in (v, t) if x.grad is not None])
                    dist.all_reduce(flat)
                if need_tf:
                    h_gtf.copy_(t.grad, non_blocking=True)
            if timed: e[3].record()
            h_loss.copy_(loss.detach().reshape(1), non_blocking=True)
            if timed: e[4].record()
            torch.cuda.synchronize()                                                  # the user reads the loss every step
            if timed:
                for name, x, y in (("wait_for_inputs", 0, 1), ("forward", 1, 2), ("loss+backward", 2, 3), ("d2h+sync", 3, 4)):
                    e2e_phase[name] += e[x].elapsed_time(e[y])
            return float(h_loss[0])

        for _ in range(max(args.warmup, 3)):
            e2e_step()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        a, b = ev(), ev()
        a.record()
        step_ms = []
        for _ in range(args.steps):
            t_s = time.perf_counter()
            e2e_step(True)
            step_ms.append(round(1e3 * (time.perf_counter() - t_s), 2))
        b.record()
        torch.cuda.synchronize()
        ms = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        e2e = {"value": all_samples * args.steps / (float(ms[0]) * 1e-3) / 1e9, "unit": "Gsamples/s",
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "ms_per_step": float(ms[0]) / args.steps,
               "phase_ms_per_step": {k: v / args.steps for k, v in e2e_phase.items()}, "wall_ms_each_step": step_ms,
               "api": "differender_b200.Raycaster.forward + loss.backward()" if mode != "nondiff" else "Raycaster.raycast_nondiff",
               "overlap": "inputs are double-buffered: step i+1's H2D copies (all inputs, every step) run on a second stream under step i's compute; "
                          "`wait_for_inputs` is what a step still waits for them"}

    # L2 -> SM read bandwidth of this box (not in MEASURED_PEAKS.json, so measured here): repeated reduction
    # of a 48 MiB buffer that stays L2-resident (126 MB L2); a library reduction, so a lower bound of the hardware figure
    l2_gbs = None
    if rank == 0:
        lb = torch.empty(48 << 18, dtype=torch.float32, device=dev).normal_()
        for _ in range(5):
            lb.sum()
        a, b = ev(), ev()
        a.record()
        for _ in range(20):
            lb.sum()
        b.record()
        torch.cuda.synchronize()
        l2_gbs = 20 * lb.numel() * 4 / (a.elapsed_time(b) * 1e-3) / 1e9
        del lb
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        hbm = float(peaks.get("hbm_gbs", 6650.0))
        vox_b = 2 if cfg["dtype"] == "f16" else 4
        s = samples_per_step[0]
        fwd_ms, bwd_ms = phase_ms["fwd"] / args.steps, phase_ms["bwd"] / args.steps
        # algorithmic bytes per sample (SURVEY 8(d) / BASELINE.md 5): forward 8 corner voxels; backward 8 corner reads +
        # 8 fp32 atomic read-modify-writes (TF-only backward = forward figure)
        fwd_bytes = 8 * vox_b
        bwd_bytes = 8 * vox_b + (64 if need_vol else 0)
        dominant = "bwd_kernel" if (mode != "nondiff" and bwd_ms >= fwd_ms) else "fwd_kernel"
        dom_bytes, dom_ms = (bwd_bytes, bwd_ms) if dominant == "bwd_kernel" else (fwd_bytes, fwd_ms)
        achieved = dom_bytes * s / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
        # DRAM bytes of the dominant kernel per launch, from the committed `ncu --set full` capture of this very workload
        traffic = None
        ncu_note = {"source": None}
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic_r01.json")))
            if tj["config"] == args.config and tj["views_per_gpu"] == views and tj["layout"] == vr.resolve_layout(vol_lin):
                traffic = tj["kernels"][dominant]["traffic_bytes_per_launch"]
                ncu_note = {"source": tj["source"],
                            "issue_active_pct": {k: round(v["issue_active_pct"], 1) for k, v in tj["kernels"].items()},
                            "l1tex_throughput_pct": {k: round(v["l1tex_throughput_pct"], 1) for k, v in tj["kernels"].items()},
                            "lts_throughput_pct": {k: round(v["lts_throughput_pct"], 1) for k, v in tj["kernels"].items()}}
        except (OSError, KeyError, ValueError):
            pass
        line = {
            "metric": METRIC[mode],
            "value": value, "unit": "Gsamples/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": cfg["desc"], "volume": f"{n}^3 {cfg['dtype']}", "image": f"{w}x{h}", "views_per_gpu": views,
                       "volume_layout": vr.resolve_layout(vol_lin), "tf": tf_name, "tf_resolution": R, "sampling_rate": sr, "max_samples": M, "jitter": cfg["jitter"],
                       "parallelism": f"views sharded over {world} GPU(s), volume+TF replicated" + (", grads all-reduced (NCCL)" if world > 1 else ""),
                       "l2": f"flushed between steps ({L2_FLUSH_BYTES >> 20} MiB memset); per-step working set also exceeds L2",
                       "active_samples_per_step_per_gpu": s, "shaded_fraction_of_active_samples": round(shaded_fraction, 4),
                       "empty_space_skipping": not args.no_skip,
                       "note": "samples whose TF alpha is exactly 0 are composited exactly without evaluating their normal, and runs of them inside "
                               "macro-cells that are transparent under the TF are counted without being marched (exact; DESIGN.md 4)"},
            "fwd": {"value": s / (fwd_ms * 1e-3) / 1e9 if fwd_ms > 0 else None, "unit": "Gsamples/s", "ms": fwd_ms},
            "bwd": {"value": s / (bwd_ms * 1e-3) / 1e9 if bwd_ms > 0 else None, "unit": "Gsamples/s", "ms": bwd_ms},
            "phase_ms_per_step": {k: v / args.steps for k, v in phase_ms.items()},
            "roofline": {"bound": "hbm", "kernel": dominant, "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm,
                         "traffic": traffic, "algorithmic_bytes_per_launch": dom_bytes * s, "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)",
                         "algorithmic_bytes_per_sample": dom_bytes,
                         "note": "algorithmic bytes are the L2-level figure of SURVEY 8(d) (8 corner reads [+ 8 fp32 atomic RMWs] per sample); the kernels are instruction-issue-bound and most accesses hit L1/L2, so measured DRAM traffic (`traffic`) is far BELOW the algorithmic bytes"},
            "rooflines_other": {"l2_read_gbs_measured": l2_gbs, "l2_how": "torch.sum over a 48 MiB L2-resident buffer, 20 reps",
                                "achieved_over_l2": (achieved / l2_gbs) if l2_gbs else None,
                                "ncu": ncu_note,
                                "binding": "instruction issue (backward 75 % issue-active, FMA pipe 56 %, ALU 41 %; forward 69 % issue-active and "
                                           "latency-exposed, L1TEX 43 %); DRAM < 2 % of peak"},
            "allreduce": None if world == 1 else {
                "bytes": int((n ** 3 + R * 4) * 4 if need_vol else R * 16), "ms": phase_ms["post"] / args.steps,
                "busbw_gbs": ((n ** 3 + R * 4) * 4 if need_vol else R * 16) * 2 * (world - 1) / world / (phase_ms["post"] / args.steps * 1e-3) / 1e9,
                "nvlink_peak_gbs": 900.0, "note": "one all_reduce(SUM) of the flat [volume grad | TF grad] fp32 buffer, rank-0 CUDA-event time of the `post` phase"},
            "clocks": clocks, "gpu_launches": launches[0],
        }
        if e2e is not None:
            line["e2e"] = e2e
        if not args.no_cpu_baseline and world == 1:
            cb = cpu_sample(cfg)
            line["cpu_baseline"] = {"value": cb["value"], "unit": "Gsamples/s", "cores": cb["cores"], "kind": "port", "sample": cb["sample"]}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    cfg = CONFIGS[args.config]
    if args.impl == "reference":
        run_reference(args, cfg)
    else:
        run_ours(args, cfg)


if __name__ == "__main__":
    main()
