"""Raw uint8 volume ingest: the reference loads `skull.raw` as
np.swapaxes(np.fromfile(path, uint8).reshape(256,256,256), 0, 1).astype(float32) / 255.0
(examples/taichi_volume_raycaster.py:548-550); here the swap, the conversion and the scaling are one kernel on the device."""
import ctypes

import numpy as np
import torch

from .. import _lib

__all__ = ["volume_from_raw_u8"]


def volume_from_raw_u8(raw, shape=None, swap_axes01=True, dtype=torch.float32, device="cuda"):
    """raw: path to a .raw file, a numpy uint8 array or a uint8 tensor with `shape` (A, B, C) (default: the array's own).
    Returns the (1, D, H, W) volume tensor the Raycaster takes, with D,H = (B, A) if swap_axes01 else (A, B), values u8/255.
    dtype=torch.uint8 keeps the bytes as they are (only the axis swap is applied): the Raycaster marches a uint8 volume
    directly (DR_VOX_U8: value u8/255 formed in registers, 8-byte cell records), bit-identical to the fp32 conversion."""
    if isinstance(raw, str):
        raw = np.fromfile(raw, dtype=np.uint8)
    t = torch.as_tensor(raw)
    if t.dtype != torch.uint8:
        raise ValueError("raw volume must be uint8")
    if shape is not None:
        t = t.reshape(shape)
    if t.ndim != 3:
        raise ValueError("raw volume must be 3-D (give `shape` for flat data)")
    t = t.to(device).contiguous()
    if not t.is_cuda:
        raise RuntimeError("volume_from_raw_u8 runs on CUDA only (no CPU fallback)")
    A, B, C = t.shape
    D, H = (B, A) if swap_axes01 else (A, B)
    if dtype == torch.uint8:
        return (t.swapaxes(0, 1) if swap_axes01 else t).contiguous()[None]
    vox = _lib.VOX_F16 if dtype == torch.float16 else _lib.VOX_F32
    d = _lib.make_desc(C, D, H, 8, 8, 2, 1, 1, 1, 1, vox, 0, 1.0, 30.0, 0.1)       # X = W, Y = D, Z = H
    out = torch.empty((1, D, H, C), dtype=dtype, device=t.device)
    with torch.cuda.device(t.device):
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        _lib.check(_lib.load().dr_ingest_u8(ctypes.byref(d), _lib.ptr(t), _lib.ptr(out), int(bool(swap_axes01)), st), "dr_ingest_u8")
    return out
