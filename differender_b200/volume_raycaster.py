"""PyTorch-facing raycaster API, mirroring the reference's differender/volume_raycaster.py.

Same names, arguments and return shapes as the reference (`Raycaster` :478-574, `RaycastFunction` :392-476,
`VolumeRaycaster` :56-116), but the Taichi kernels are replaced by the sm_100a library behind include/diffrender.h.
There is no Taichi, no Triton, no CPU fallback: non-CUDA tensors raise.

Differences from the reference, all additive or fixes listed in SURVEY.md 7.3:
  * `forward(..., jitter_tensor=None)`: jitter is a supplied uniform [0,1) tensor ([BS,]H,W) in output-image orientation
    (drawn with torch.rand when `jitter=True` and none is given) and the SAME tensor is used by the backward (H7).
  * un-batched volume / TF with batched cameras are shared, not cloned BS times (:566-567), and their gradient is
    accumulated once on the device instead of materialising BS per-view gradients.
  * fp16 volumes stay fp16 in HBM (no `.float()` up-cast, :119); arithmetic is fp32.
  * the forward keeps no state in the object (:429-430): everything the backward needs is in `ctx` (H10).
"""
import ctypes

import torch

from . import _lib
from ._lib import (F_HAS_JITTER, F_LAYOUT_BRICK8, F_LAYOUT_CELL8, F_NEEDS_TF_GRAD, F_NEEDS_VOL_GRAD, F_NONDIFF, F_OUT_IMAGE, F_TF_4R, VOX_F16,
                   VOX_F32, VOX_U8)

__all__ = ["VolumeRaycaster", "RaycastFunction", "RaycastMSEFunction", "Raycaster"]


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


VOLUME_DTYPES = (torch.float32, torch.float16, torch.uint8)      # how a volume may be STORED; arithmetic is always fp32


def _vox(dtype):
    """DR_VOX_* of a stored volume dtype.  uint8 voxels mean u8 / 255 (the reference's skull.raw ingest,
    examples/taichi_volume_raycaster.py:548-550) and are marched from an 8-byte-per-cell copy without ever being widened in HBM."""
    return VOX_F16 if dtype == torch.float16 else (VOX_U8 if dtype == torch.uint8 else VOX_F32)


class VolumeRaycaster:
    """Replaces the reference's Taichi-side object (:56-116): resolutions and camera constants only.

    It owns no device memory and no global runtime (the reference calls ti.init, :486, and allocates a
    16*W*H*max_samples-byte render tape plus its gradient, :82,102-103).
    """

    def __init__(self, volume_resolution, render_resolution, max_samples=512, tf_resolution=128, fov=30.0,
                 nearfar=(0.1, 100.0), layout="auto", skip_empty=True):
        if layout not in ("auto", "linear", "brick8", "cell8"):
            raise ValueError("layout must be 'auto', 'linear' (read the torch tensor in place), 'brick8' (8x8x8-bricked copy) "
                             "or 'cell8' (cell-major copy, 8x the volume's bytes)")
        self.layout = layout
        self.skip_empty = bool(skip_empty)   # exact empty-space skipping in the forward march (dr_build_skip_grid / dr_forward_ex)
        self._skip_ring, self._skip_pending, self._skip_use = None, [], True     # asynchronous read-back of the grids' empty counts
        self._skip_calls, self._skip_minmax, self._copy_cache, self._auto_layout, self._skip_stream = 0, None, None, {}, None
        self.volume_resolution = tuple(int(v) for v in volume_resolution)     # Taichi order (X, Y, Z) = torch (W, D, H)
        self.resolution = tuple(int(v) for v in render_resolution)            # (w, h)
        self.max_samples = int(max_samples)
        self.tf_resolution = int(tf_resolution)
        self.fov_deg = float(fov)
        self.near, self.far = float(nearfar[0]), float(nearfar[1])
        self.ambient, self.diffuse, self.specular, self.shininess = 0.4, 0.8, 0.3, 32.0    # :91-94
        self.last_K = None           # per-ray active sample counts of the most recent forward ([BS,H,W] int32)
        # FusedVolumeSGD (optim.py): the autograd backward leaves the volume gradient un-gathered in `pending_grad_cells`
        # (cell-major, accumulated over backward calls) and returns None for the volume; dr_gather_step consumes it
        self.defer_volume_gather = False
        self.pending_grad_cells = None
        # DistributedRaycaster: {"vol": [1,Y,Z,X] fp32 view of its flat all-reduce buffer}; the gather writes there directly
        self.grad_sink = None
        self.last_skip_grid = None
        self.kernel_launches = 0     # this object's kernel launches so far (the library's own kernels only; bench.py reports the count)

    @property
    def max_valid_sample_step_count(self):
        """Largest number of active samples on any ray of the most recent forward -- the reference's diagnostic of the same
        name (:89, :370-372; printed as "Max Samples: x / M" by its demo).  Synchronises."""
        return 0 if self.last_K is None else int(self.last_K.max().item())

    # -- thin wrappers over the C ABI (one distinct volume is bricked once, not once per view) ---------------
    def desc(self, BS, Bvol, Btf, vox_dtype, flags, sampling_rate):
        X, Y, Z = self.volume_resolution
        w, h = self.resolution
        return _lib.make_desc(X, Y, Z, w, h, self.tf_resolution, self.max_samples, BS, Bvol, Btf, vox_dtype, flags,
                              sampling_rate, self.fov_deg, self.near)

    # Measured on B200 (profiles/r01_experiments.md): the cell-major copy wins at every benchmark size
    # (forward +18 % at 256^3, +61 % at 512^3, +48 % at 1024^3 fp16 over the previous best layout; backward +0..21 %) because a
    # cell is one sector and two 16-byte loads instead of eight 4-byte loads spread over ~5 sectors each.  It costs 8x the
    # volume's bytes, so `auto` falls back to the 8x8x8-bricked copy (1x) above AUTO_CELL_BYTES of cell-major data.
    AUTO_CELL_BYTES = 48 << 30
    AUTO_FREE_FRACTION = 0.8                     # of the memory that is free right now (driver-free + the allocator's idle blocks)

    @staticmethod
    def _free_bytes(device):
        """Bytes a new allocation can get on `device` without an out-of-memory error: what the driver reports free plus
        the blocks the caching allocator holds but has not handed out."""
        free, _ = torch.cuda.mem_get_info(device)
        return free + torch.cuda.memory_reserved(device) - torch.cuda.memory_allocated(device)

    def resolve_layout(self, vol_lin, need_vol_grad=False):
        """`auto`: the cell-major copy (8x the volume's bytes) when it is at most AUTO_CELL_BYTES AND it fits -- together with
        the 32-byte-per-voxel cell-major gradient buffer when the volume gradient is wanted -- into AUTO_FREE_FRACTION of
        the device memory that is free now; otherwise the 8x8x8-bricked copy (1x)."""
        X, Y, Z = self.volume_resolution
        if vol_lin.dtype == torch.uint8:
            if self.layout not in ("auto", "cell8") or max(X, Y, Z) > 2000:
                raise ValueError("uint8 volumes are marched from the cell-major copy (layout 'auto' or 'cell8', axes <= 2000 voxels); "
                                 "convert with volume_from_raw_u8(dtype=torch.float16 / float32) for the other layouts")
            return "cell8"
        if self.layout != "auto":
            return self.layout
        if max(X, Y, Z) > 2000:
            return "linear"                      # the generic tap path exists for the linear layout only
        cached = self._cached_copy(vol_lin)
        if cached is not None:
            return "cell8" if cached.ndim == 3 else "brick8"
        # decided once per (batch, dtype, gradient wanted): cudaMemGetInfo costs 0.2-4 ms of host time per call on B200
        # (measured as a bubble in front of every forward), and a loop should not flip layouts from step to step
        key = (vol_lin.shape[0], vol_lin.dtype, bool(need_vol_grad), str(vol_lin.device))
        if key not in self._auto_layout:
            copy = X * Y * Z * 8 * vol_lin.element_size() * vol_lin.shape[0]
            need = copy + (X * Y * Z * 32 * vol_lin.shape[0] if need_vol_grad else 0)
            fits = copy <= self.AUTO_CELL_BYTES and not (vol_lin.is_cuda and need > self.AUTO_FREE_FRACTION * self._free_bytes(vol_lin.device))
            self._auto_layout[key] = "cell8" if fits else "brick8"
        return self._auto_layout[key]

    # -- caches keyed on a source tensor ------------------------------------------------------------------------------
    # A key is (storage address, data pointer, version counter, shape, dtype).  Every cache entry also HOLDS the source's
    # storage: while the entry lives the caching allocator cannot hand the same address to another tensor, so a fresh
    # tensor (version 0) with different contents can never match a stale key.  In-place edits through PyTorch bump the
    # version counter; kernels of this library that write a tensor behind PyTorch's back bump it themselves
    # (torch.autograd.graph.increment_version).
    @staticmethod
    def _src_key(t):
        return (t.untyped_storage().data_ptr(), t.data_ptr(), t._version, tuple(t.shape), t.dtype)

    def _cached_copy(self, vol_lin):
        c = self._copy_cache
        return c[2] if c is not None and c[0] == self._src_key(vol_lin) else None

    def _lflag(self, vol):
        """Layout flag for a tensor returned by brick(): a 2-D tensor is a bricked copy, a 3-D one [Bvol, cells, 8] the
        cell-major copy, a 4-D one the linear volume."""
        return F_LAYOUT_BRICK8 if vol.ndim == 2 else (F_LAYOUT_CELL8 if vol.ndim == 3 else 0)

    def brick(self, vol_lin, need_vol_grad=False):
        """[Bvol, Y, Z, X] contiguous fp32/fp16 CUDA tensor -> what the march kernels read: the tensor itself for the
        'linear' layout (zero copy), a bricked copy [Bvol, elems] for 'brick8', or the cell-major copy [Bvol, X*Y*Z, 8]
        for 'cell8' (same dtype).  The copy is cached while the same tensor is unchanged (one entry; forget_volume()
        drops it), so a loop that only changes the transfer function or the cameras re-lays the volume once."""
        X, Y, Z = self.volume_resolution
        if tuple(vol_lin.shape[1:]) != (Y, Z, X):
            raise ValueError(f"volume has spatial shape {tuple(vol_lin.shape[1:])}, raycaster was built for (D,H,W)={(Y, Z, X)}")
        if not vol_lin.is_contiguous():
            raise ValueError("volume must be contiguous")
        layout = self.resolve_layout(vol_lin, need_vol_grad)
        if layout == "linear":
            return vol_lin
        cached = self._cached_copy(vol_lin)
        if cached is not None and self._lflag(cached) == (F_LAYOUT_CELL8 if layout == "cell8" else F_LAYOUT_BRICK8):
            return cached
        self._copy_cache = None                  # release the previous copy before allocating the next one
        # (the copies below remember the linear tensor they were made from: the forward builds its skip grid from it)
        vox = _vox(vol_lin.dtype)
        d = self.desc(1, 1, 1, vox, 0, 1.0)
        d.Bvol = vol_lin.shape[0]
        lib = _lib.load()
        if layout == "cell8":
            out = torch.empty((vol_lin.shape[0], X * Y * Z, 8), dtype=vol_lin.dtype, device=vol_lin.device)
            _lib.check(lib.dr_expand_cells(ctypes.byref(d), _lib.ptr(vol_lin), _lib.ptr(out), _stream()), "dr_expand_cells")
            self.kernel_launches += 1
        else:
            out = torch.empty((vol_lin.shape[0], lib.dr_bricked_elems(ctypes.byref(d))), dtype=vol_lin.dtype, device=vol_lin.device)
            _lib.check(lib.dr_brick_volume(ctypes.byref(d), _lib.ptr(vol_lin), _lib.ptr(out), _stream()), "dr_brick_volume")
            self.kernel_launches += 1
        out.dr_source = vol_lin
        self._copy_cache = (self._src_key(vol_lin), vol_lin.untyped_storage(), out)
        return out

    def forget_volume(self):
        """Drops the cached volume copy and per-macro-cell min / max (and the references that keep their source volume alive).
        The caches are keyed on the volume tensor's storage, pointer and version counter, so this is only needed when the memory
        was changed behind PyTorch's back, to release the memory -- or by a benchmark that wants them rebuilt every step."""
        self._skip_minmax = None
        self._copy_cache = None

    def skip_grid(self, d, bricked, tf_r4):
        """Macro-cell emptiness bytes for this call (exact empty-space skipping), or None when skipping is off, the volume's
        linear tensor is not known (a copy not made by brick()) or the generic tap path is in use."""
        src = bricked if bricked.ndim == 4 else getattr(bricked, "dr_source", None)
        if not self.skip_empty or src is None or d.tap_generic:
            return None
        self._skip_calls += 1
        if not self._skip_use and self._skip_calls % self.SKIP_RETRY:
            return None                          # not worth it lately: look again every SKIP_RETRY-th call
        lib = _lib.load()
        # the per-macro-cell min / max depends on the volume only: reused while the same tensor has not been written to
        # (the entry holds the volume's storage, so its address cannot be recycled for another volume while the key lives)
        key = self._src_key(src)
        mm_valid = self._skip_minmax is not None and self._skip_minmax[0] == key
        mm = self._skip_minmax[2] if mm_valid else \
            torch.empty(max(lib.dr_skip_minmax_bytes(ctypes.byref(d)) // 4, 2), dtype=torch.float32, device=src.device)
        grid = torch.empty(max(lib.dr_skip_grid_bytes(ctypes.byref(d)), 1), dtype=torch.uint8, device=src.device)
        _lib.check(lib.dr_build_skip_grid(ctypes.byref(d), _lib.ptr(src), _lib.ptr(tf_r4), _lib.ptr(mm), int(mm_valid), _lib.ptr(grid),
                                          _stream()), "dr_build_skip_grid")
        self.kernel_launches += 1 if mm_valid else 2             # skip_classify_kernel (+ skip_minmax_kernel)
        self._skip_minmax = (key, src.untyped_storage(), mm)
        # Performance hint only (results are bit-identical either way): with (almost) no empty macro-cells the skip kernels'
        # bookkeeping costs ~5 % of the forward, so they are not used while the PREVIOUS call's grid -- its count is read back
        # asynchronously, never waited for -- had fewer than MIN_EMPTY_FRACTION of its macro-cells empty.
        if self._skip_ring is None:
            self._skip_ring = torch.zeros(self.SKIP_RING, dtype=torch.int32).pin_memory()
        while self._skip_pending and self._skip_pending[0][1].query():           # newest completed read-back decides
            slot, _, total = self._skip_pending.pop(0)
            self._skip_use = int(self._skip_ring[slot]) >= self.MIN_EMPTY_FRACTION * total
        if len(self._skip_pending) < self.SKIP_RING:                             # (a full ring just skips this call's read-back)
            used = {p[0] for p in self._skip_pending}
            slot = next(i for i in range(self.SKIP_RING) if i not in used)
            # The 4-byte read-back runs on a SIDE stream: on the march's own stream it would sit between the classification and the
            # forward kernel, and a device-to-host copy -- however small -- queues behind whatever else occupies the copy engine
            # (measured: a 64 MiB gradient download in flight delayed every forward by 0.9 ms on one GPU and by 8 ms on a busy 8-GPU host).
            main = torch.cuda.current_stream()
            if self._skip_stream is None or self._skip_stream.device != src.device:
                self._skip_stream = torch.cuda.Stream(device=src.device)
            side = self._skip_stream
            side.wait_stream(main)
            with torch.cuda.stream(side):
                self._skip_ring[slot:slot + 1].copy_(grid[:4].view(torch.int32), non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(side)
            grid.record_stream(side)
            self._skip_pending.append((slot, ev, grid.numel() - 16))
        return grid if self._skip_use else None

    MIN_EMPTY_FRACTION = 0.05
    SKIP_RING = 8
    SKIP_RETRY = 16

    def march(self, bricked, tf_r4, cam, sampling_rate, jitter=None, nondiff=False, image_layout=True, want_aux=True,
              extra_flags=0, mse_target=None, skip=None):
        """Forward of cam.shape[0] views.  Returns (out, K, Tprev), or (out, K, Tprev, loss_sum[BS]) when `mse_target`
        (same layout as out) is given: the squared-error sum is accumulated in the kernel's epilogue.
        `tf_r4` is [Btf, R, 4] (the reference's order); pass extra_flags=F_TF_4R with a [Btf, 4, R] tensor (torch order).
        `skip` = False marches every sample; None / True use the exact empty-space skip grid when `skip_empty` is set."""
        BS = cam.shape[0]
        if extra_flags & F_TF_4R:
            tf_r4 = _Tf4R(tf_r4)
        w, h = self.resolution
        vox = _vox(bricked.dtype)
        flags = (F_NONDIFF if nondiff else 0) | (F_HAS_JITTER if jitter is not None else 0) | (F_OUT_IMAGE if image_layout else 0) | \
                self._lflag(bricked) | extra_flags
        d = self.desc(BS, bricked.shape[0], tf_r4.shape[0], vox, flags, sampling_rate)
        dev = bricked.device
        out = torch.empty((BS, 4, h, w) if image_layout else (BS, w, h, 4), dtype=torch.float32, device=dev)
        K = torch.empty((BS, h, w), dtype=torch.int32, device=dev) if want_aux else None
        Tp = torch.empty((BS, h, w), dtype=torch.float32, device=dev) if (want_aux and not nondiff) else None
        grid = self.skip_grid(d, bricked, tf_r4) if skip is None or skip else None
        self.last_skip_grid = grid           # the volume-only backward of the same call can jump over the same empty runs
        loss_sum = torch.zeros((BS,), dtype=torch.float32, device=dev) if mse_target is not None else None
        _lib.check(_lib.load().dr_forward_ex(ctypes.byref(d), _lib.ptr(bricked), _lib.ptr(tf_r4), _lib.ptr(cam), _lib.ptr(jitter),
                                             _lib.ptr(mse_target), _lib.ptr(grid), _lib.ptr(out), _lib.ptr(K), _lib.ptr(Tp),
                                             _lib.ptr(loss_sum), _stream()), "dr_forward_ex")
        self.kernel_launches += 1
        if mse_target is not None:
            return out, K, Tp, loss_sum
        return out, K, Tp

    def march_backward(self, bricked, tf_r4, cam, sampling_rate, jitter, grad_out, out, K, Tprev, need_vol, need_tf,
                       image_layout=True, grad_cells=None, extra_flags=0, mse_scale=None, skip_grid=None, mse_scale_dev=None):
        """Backward of cam.shape[0] views.  Returns (grad_vol_linear [Bvol,Y,Z,X] fp32 or None, grad_tf [Btf,R,4] or None).
        The volume gradient is scattered into a cell-major buffer [Bvol, X*Y*Z*8] (zeroed here) and gathered once.
        If `grad_cells` is given it is accumulated into and NOT gathered (it is returned instead), so that several calls
        (e.g. chunks of a large view batch) share one buffer and one gather.
        With `mse_scale`, `grad_out` is the TARGET image and dL/d(out) = mse_scale*(out - target) is formed in the kernel
        (times the fp32 device scalar `mse_scale_dev` when given: an upstream gradient that never has to visit the host).
        `skip_grid`: the grid the forward of the same (volume, TF) built (march() leaves it in `last_skip_grid`); only the
        volume-only backward (need_tf False) uses it, to jump over runs of samples in exactly transparent macro-cells."""
        BS = cam.shape[0]
        if extra_flags & F_TF_4R:
            tf_r4 = _Tf4R(tf_r4)
        vox = _vox(bricked.dtype)
        flags = (F_HAS_JITTER if jitter is not None else 0) | (F_OUT_IMAGE if image_layout else 0) | \
                (F_NEEDS_VOL_GRAD if need_vol else 0) | (F_NEEDS_TF_GRAD if need_tf else 0) | extra_flags | self._lflag(bricked)
        d = self.desc(BS, bricked.shape[0], tf_r4.shape[0], vox, flags, sampling_rate)
        dev = bricked.device
        lib = _lib.load()
        keep_cells = grad_cells is not None
        if need_vol and grad_cells is None and self.defer_volume_gather and bricked.shape[0] == 1:
            if self.pending_grad_cells is None:
                self.pending_grad_cells = torch.zeros((1, lib.dr_grad_cells_elems(ctypes.byref(d))), dtype=torch.float32, device=dev)
            grad_cells, keep_cells = self.pending_grad_cells, True
        if need_vol and grad_cells is None:
            grad_cells = torch.zeros((bricked.shape[0], lib.dr_grad_cells_elems(ctypes.byref(d))), dtype=torch.float32, device=dev)
        gtf = torch.zeros(tf_r4.t.shape if isinstance(tf_r4, _Tf4R) else tf_r4.shape, dtype=torch.float32, device=dev) if need_tf else None
        ws_bytes = lib.dr_workspace_bytes(ctypes.byref(d))
        ws = torch.empty((max(ws_bytes, 16) + 3) // 4, dtype=torch.float32, device=dev)
        fused = mse_scale is not None
        _lib.check(lib.dr_backward_ex(ctypes.byref(d), _lib.ptr(bricked), _lib.ptr(tf_r4), _lib.ptr(cam), _lib.ptr(jitter),
                                      None if fused else _lib.ptr(grad_out), _lib.ptr(grad_out) if fused else None,
                                      ctypes.c_float(mse_scale if fused else 0.0), _lib.ptr(mse_scale_dev) if fused else None,
                                      _lib.ptr(skip_grid) if (need_vol and not need_tf) else None,
                                      _lib.ptr(out), _lib.ptr(K), _lib.ptr(Tprev), _lib.ptr(grad_cells) if need_vol else None, _lib.ptr(gtf),
                                      _lib.ptr(ws), ws_bytes, _stream()), "dr_backward_ex")
        self.kernel_launches += 1 + (1 if need_tf else 0)           # bwd_kernel (+ tf_reduce_kernel)
        if not need_vol:
            return None, gtf
        if keep_cells:
            return grad_cells, gtf
        sink = self.grad_sink.get("vol") if self.grad_sink else None
        if sink is not None and (grad_cells.shape[0] != 1 or sink.device != grad_cells.device):
            sink = None
        return self.gather(grad_cells, out=sink), gtf

    def gather(self, grad_cells, out=None):
        """Cell-major gradient [Bvol, X*Y*Z*8] -> linear [Bvol, Y, Z, X] fp32 with nan_to_num (into `out` if given, e.g. a
        slice of the flat buffer that is all-reduced across GPUs)."""
        X, Y, Z = self.volume_resolution
        d = self.desc(1, 1, 1, VOX_F32, 0, 1.0)
        d.Bvol = grad_cells.shape[0]
        gl = out if out is not None else torch.empty((grad_cells.shape[0], Y, Z, X), dtype=torch.float32, device=grad_cells.device)
        if tuple(gl.shape) != (grad_cells.shape[0], Y, Z, X) or gl.dtype != torch.float32 or not gl.is_contiguous():
            raise ValueError("gather: `out` must be a contiguous fp32 tensor of shape [Bvol, Y, Z, X]")
        _lib.check(_lib.load().dr_gather_grad(ctypes.byref(d), _lib.ptr(grad_cells), _lib.ptr(gl), 0, _stream()), "dr_gather_grad")
        self.kernel_launches += 1
        return gl


class _Tf4R:
    """Adapter so that a [Btf, 4, R] transfer function reports the [Btf, R, 4] shape the wrappers size their buffers from."""

    def __init__(self, t):
        self.t = t
        self.shape = (t.shape[0], t.shape[2], t.shape[1])

    def data_ptr(self):
        return self.t.data_ptr()


def _require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("differender_b200: the ray-march runs on CUDA (sm_100a) only; got a CPU tensor. "
                               "There is no CPU fallback.")


def _prepare(vr, volume, tf, look_from, batched, jitter, jitter_tensor):
    """Normalises the reference-style inputs (:395-410) to what the C ABI takes, without cloning shared inputs.
    Returns (BS, vol_b, vol_lin [Bvol,Y,Z,X], tf_r4 [Btf,R,4], cam [BS,3], jit [BS,H,W] or None)."""
    _require_cuda(volume, tf, look_from, jitter_tensor)
    is_batched, bs = batched
    BS = int(bs) if is_batched else 1
    w, h = vr.resolution
    vol_b = volume.ndim == 4
    v = volume if vol_b else volume[None]
    if vol_b and v.shape[0] > 1 and v.stride(0) == 0:
        v = v[:1]                                          # an .expand()ed shared volume: do not clone it (:566)
    if v.dtype not in VOLUME_DTYPES:
        v = v.float()                                      # set_volume's .float() (:119); fp16 and uint8 (= u8 / 255) are kept as stored
    vol_lin = v.permute(0, 2, 3, 1).contiguous()           # [Bvol, Y, Z, X] == torch (D, H, W); no-op for our views
    if vol_lin.shape[0] not in (1, BS):
        raise ValueError(f"volume batch {vol_lin.shape[0]} does not match batch size {BS}")
    t = tf if tf.ndim == 3 else tf[None]
    if t.shape[0] > 1 and t.stride(0) == 0:
        t = t[:1]
    tf_r4 = t.float().contiguous()                         # [Btf, R, 4]
    if tf_r4.shape[-1] != 4 or tf_r4.shape[-2] != vr.tf_resolution:
        raise ValueError(f"tf has shape {tuple(tf.shape)}, expected ([BS,] {vr.tf_resolution}, 4)")
    if tf_r4.shape[0] not in (1, BS):
        raise ValueError(f"tf batch {tf_r4.shape[0]} does not match batch size {BS}")
    cam = look_from.float().reshape(-1, 3)
    if cam.shape[0] == 1 and BS > 1:
        cam = cam.expand(BS, 3)
    cam = cam.contiguous()
    if cam.shape[0] != BS:
        raise ValueError(f"look_from batch {cam.shape[0]} does not match batch size {BS}")
    jit = None
    if jitter:
        if jitter_tensor is None:
            jit = torch.rand((BS, h, w), dtype=torch.float32, device=volume.device)     # replaces ti.random (:255)
        else:
            jit = jitter_tensor.float().reshape(-1, h, w)
            if jit.shape[0] == 1 and BS > 1:
                jit = jit.expand(BS, h, w)
            jit = jit.contiguous()
            if jit.shape[0] != BS:
                raise ValueError(f"jitter_tensor batch {jit.shape[0]} does not match batch size {BS}")
    return BS, vol_b, vol_lin, tf_r4, cam, jit


def _save(ctx, vr, volume, tf, sampling_rate, batched, vol_b, bricked, tf_r4, cam, jit, out, K, Tp, image_layout, target=None):
    vr.last_K = K
    ctx.vr, ctx.sampling_rate, ctx.image_layout = vr, sampling_rate, image_layout
    ctx.is_batched, ctx.vol_batched, ctx.tf_batched = batched[0], vol_b, tf.ndim == 3
    ctx.vol_shape, ctx.tf_shape = tuple(volume.shape), tuple(tf.shape)
    # Everything the backward re-marches with goes through save_for_backward: tf_r4 / cam / jit alias the caller's tensors
    # when those are already fp32 and contiguous, and the zero-copy layout reads the caller's volume in place -- autograd's
    # version check then turns an in-place edit between forward and backward (tf.clamp_(), an optimiser step ...) into an
    # error instead of a backward that silently re-marches with other inputs than the forward's K / Tprev / image
    # (the reference snapshots its inputs in Taichi fields, :419-421).  `out` is saved the same way: for an output of the
    # Function save_for_backward keeps no reference cycle (a plain ctx attribute would: out.grad_fn -> ctx -> out left
    # every buffer of the step to Python's cyclic collector and cost ~1 GiB of fresh cudaMalloc per step).
    zero_copy = bricked.data_ptr() == volume.data_ptr()
    ctx.save_for_backward(volume if zero_copy else None, tf_r4, cam, jit, out, K, Tp, target, vr.last_skip_grid)
    ctx.bricked = None if zero_copy else bricked       # our own copy (cell-major / bricked layout, or a cast / contiguous copy)


def _saved(ctx):
    """(volume as the march kernels read it, tf_r4, cam, jit, out, K, Tp, mse target or None) of the forward."""
    volume, tf_r4, cam, jit, out, K, Tp, target, ctx.skip_grid = ctx.saved_tensors
    if ctx.bricked is not None:
        vol = ctx.bricked
    else:
        v = volume if ctx.vol_batched else volume[None]
        if ctx.vol_batched and v.shape[0] > 1 and v.stride(0) == 0:
            v = v[:1]
        vol = v.permute(0, 2, 3, 1).contiguous()           # the same zero-copy view the forward read
    return vol, tf_r4, cam, jit, out, K, Tp, target


def _shape_grads(ctx, gvol, gtf, need_vol, need_tf):
    gv = gt = None
    if need_vol and gvol is ctx.vr.pending_grad_cells:
        need_vol = False                                       # deferred: FusedVolumeSGD gathers and applies it (volume.grad stays None)
    if need_vol:
        gv = gvol.permute(0, 3, 1, 2)                          # [Bvol, X, Y, Z] view (Taichi order, :447)
        if not ctx.vol_batched:
            gv = gv[0]
        elif gv.shape[0] != ctx.vol_shape[0]:
            gv = gv.expand(ctx.vol_shape)                      # caller passed an expanded shared volume
    if need_tf:
        gt = gtf if ctx.tf_batched else gtf[0]
        if ctx.tf_batched and gt.shape[0] != ctx.tf_shape[0]:
            gt = gt.expand(ctx.tf_shape)
    return gv, gt


class RaycastFunction(torch.autograd.Function):
    """Drop-in for the reference's autograd Function (:392-476); same positional arguments, two optional extras."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=None)
    def forward(ctx, vr, volume, tf, look_from, sampling_rate, batched, jitter=True, jitter_tensor=None, image_layout=False):
        """volume ([BS,] X, Y, Z) in the reference's Taichi order (a permuted view of the torch tensor is fine, no copy is
        made when the underlying memory is the contiguous torch (D,H,W) tensor); tf ([BS,] R, 4); look_from ([BS,] 3).
        Returns ([BS,] W, H, 4) like the reference, or ([BS,] 4, H, W) already flipped when image_layout=True."""
        _require_cuda(volume, tf, look_from, jitter_tensor)
        with torch.cuda.device(volume.device):
            BS, vol_b, vol_lin, tf_r4, cam, jit = _prepare(vr, volume, tf, look_from, batched, jitter, jitter_tensor)
            bricked = vr.brick(vol_lin, need_vol_grad=ctx.needs_input_grad[1])
            out, K, Tp = vr.march(bricked, tf_r4, cam, sampling_rate, jit, nondiff=False, image_layout=image_layout)
        _save(ctx, vr, volume, tf, sampling_rate, batched, vol_b, bricked, tf_r4, cam, jit, out, K, Tp, image_layout)
        return out if batched[0] else out[0]

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, grad_output):
        need_vol, need_tf = ctx.needs_input_grad[1], ctx.needs_input_grad[2]
        if not (need_vol or need_tf):
            return (None,) * 9
        with torch.cuda.device(grad_output.device):
            go = grad_output if ctx.is_batched else grad_output[None]
            go = go.float().contiguous()
            vol, tf_r4, cam, jit, out, K, Tp, _ = _saved(ctx)
            gvol, gtf = ctx.vr.march_backward(vol, tf_r4, cam, ctx.sampling_rate, jit, go, out, K, Tp, need_vol, need_tf,
                                              image_layout=ctx.image_layout, skip_grid=ctx.skip_grid)
        gv, gt = _shape_grads(ctx, gvol, gtf, need_vol, need_tf)
        return None, gv, gt, None, None, None, None, None, None


class RaycastMSEFunction(torch.autograd.Function):
    """Render + mean-squared-error against `target` in one pass each way: the loss is reduced in the
    forward kernel's epilogue and dL/d(image) = 2 (image - target) / numel is formed inside the backward kernel, so
    neither the residual nor the gradient image exists in HBM.  Equivalent to
    `F.mse_loss(RaycastFunction.apply(..., image_layout=True), target)` (reference examples/test_opt_tf.py:70-72)."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=None)
    def forward(ctx, vr, volume, tf, look_from, sampling_rate, batched, jitter, jitter_tensor, target):
        _require_cuda(volume, tf, look_from, jitter_tensor, target)
        with torch.cuda.device(volume.device):
            BS, vol_b, vol_lin, tf_r4, cam, jit = _prepare(vr, volume, tf, look_from, batched, jitter, jitter_tensor)
            w, h = vr.resolution
            tgt = target.float().reshape(BS, 4, h, w).contiguous()
            bricked = vr.brick(vol_lin, need_vol_grad=ctx.needs_input_grad[1])
            out, K, Tp, loss_sum = vr.march(bricked, tf_r4, cam, sampling_rate, jit, image_layout=True, mse_target=tgt)
        _save(ctx, vr, volume, tf, sampling_rate, batched, vol_b, bricked, tf_r4, cam, jit, out, K, Tp, True, target=tgt)
        img = out if batched[0] else out[0]
        ctx.mark_non_differentiable(img)
        return loss_sum.sum() / out.numel(), img

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, grad_loss, _grad_img):
        need_vol, need_tf = ctx.needs_input_grad[1], ctx.needs_input_grad[2]
        if not (need_vol or need_tf):
            return (None,) * 9
        with torch.cuda.device(grad_loss.device):
            vol, tf_r4, cam, jit, out, K, Tp, tgt = _saved(ctx)
            # the upstream gradient of the loss stays on the device (the kernel multiplies it in): float(grad_loss) would make the
            # host wait for the whole forward and leave the device idle while the backward is being launched
            up = grad_loss.detach().to(torch.float32).reshape(1).contiguous()
            gvol, gtf = ctx.vr.march_backward(vol, tf_r4, cam, ctx.sampling_rate, jit, tgt, out, K, Tp, need_vol, need_tf,
                                              image_layout=True, mse_scale=2.0 / out.numel(), skip_grid=ctx.skip_grid, mse_scale_dev=up)
        gv, gt = _shape_grads(ctx, gvol, gtf, need_vol, need_tf)
        return None, gv, gt, None, None, None, None, None, None


class Raycaster(torch.nn.Module):
    """Same constructor and methods as the reference's `Raycaster` (:478-574)."""

    def __init__(self, volume_shape, output_shape, tf_shape, sampling_rate=1.0, jitter=True, max_samples=512, fov=30.0,
                 near=0.1, far=100.0, ti_kwargs={}, layout="auto", skip_empty=True):
        super().__init__()
        self.volume_shape = (volume_shape[2], volume_shape[0], volume_shape[1])       # torch (D,H,W) -> Taichi (W,D,H) :481
        self.output_shape = output_shape
        self.tf_shape = tf_shape
        self.sampling_rate = sampling_rate
        self.jitter = jitter
        self.ti_kwargs = dict(ti_kwargs)      # accepted for signature compatibility; there is no Taichi runtime to configure
        _lib.load()                           # fail loudly at construction if the CUDA library is missing
        self.vr = VolumeRaycaster(self.volume_shape, output_shape, max_samples=max_samples, tf_resolution=tf_shape,
                                  fov=fov, nearfar=(near, far), layout=layout, skip_empty=skip_empty)

    def raycast_nondiff(self, volume, tf, look_from, sampling_rate=None):
        """Non-differentiable render (:490-523): alpha-skip, no shading clamp, output clamped to 1, jitter off,
        default sampling rate 4x."""
        with torch.no_grad(), torch.autocast("cuda", enabled=False):
            _require_cuda(volume, tf, look_from)
            batched, bs, vol_in, tf_in, lf_in = self._determine_batch(volume, tf, look_from)
            sr = sampling_rate if sampling_rate is not None else 4.0 * self.sampling_rate
            BS = bs if batched else 1
            with torch.cuda.device(volume.device):
                v = vol_in if vol_in.ndim == 4 else vol_in[None]
                if v.dtype not in VOLUME_DTYPES:
                    v = v.float()
                vol_lin = v.permute(0, 2, 3, 1).contiguous()
                t = (tf_in if tf_in.ndim == 3 else tf_in[None]).float().contiguous()
                cam = lf_in.float().reshape(-1, 3)
                if cam.shape[0] == 1 and BS > 1:
                    cam = cam.expand(BS, 3)
                cam = cam.contiguous()
                out, K, _ = self.vr.march(self.vr.brick(vol_lin), t, cam, sr, None, nondiff=True, image_layout=True)
            self.vr.last_K = K
            return out if batched else out[0]

    def forward(self, volume, tf, look_from, jitter_tensor=None):
        """volume ([BS,] 1, D, H, W), tf ([BS,] 4, R), look_from ([BS,] 3)  ->  ([BS,] 4, H, W)   (:525-548).
        The volume lives in [-1,1]^3 centred at the origin; the camera looks at the origin."""
        batched, bs, vol_in, tf_in, lf_in = self._determine_batch(volume, tf, look_from)
        return RaycastFunction.apply(self.vr, vol_in, tf_in, lf_in, self.sampling_rate, (batched, bs), self.jitter,
                                     jitter_tensor, True)

    def mse_loss(self, volume, tf, look_from, target, jitter_tensor=None):
        """Fused render + MSE: returns (loss, image) with `loss == F.mse_loss(self(volume, tf, look_from), target)`;
        the image is returned detached (gradients flow through the loss only)."""
        batched, bs, vol_in, tf_in, lf_in = self._determine_batch(volume, tf, look_from)
        return RaycastMSEFunction.apply(self.vr, vol_in, tf_in, lf_in, self.sampling_rate, (batched, bs), self.jitter,
                                        jitter_tensor, target)

    def _determine_batch(self, volume, tf, look_from):
        """Same rule as the reference (:551-571): anything batched => batch size from the first batched input.
        Returns Taichi-order VIEWS; un-batched inputs are shared, not expanded and cloned (:566-568)."""
        b_vol, b_tf, b_lf = volume.ndim == 5, tf.ndim == 3, look_from.ndim == 2
        if b_vol or b_tf or b_lf:
            bs = [volume, tf, look_from][[b_vol, b_tf, b_lf].index(True)].size(0)
            vol_out = volume.squeeze(1).permute(0, 3, 1, 2) if b_vol else volume.squeeze(0).permute(2, 0, 1)
            tf_out = tf.permute(0, 2, 1) if b_tf else tf.permute(1, 0)
            return True, bs, vol_out, tf_out, look_from
        return False, 0, volume.squeeze(0).permute(2, 0, 1), tf.permute(1, 0), look_from

    def extra_repr(self):
        return f'Volume ({self.volume_shape}), Output Render ({self.output_shape}), TF ({self.tf_shape}), Max Samples = {self.vr.max_samples}'
