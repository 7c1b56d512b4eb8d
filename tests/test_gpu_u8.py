"""uint8-stored volumes (DR_VOX_U8; SURVEY 8(f) row 4, reference examples/taichi_volume_raycaster.py:548-550: skull.raw is
256^3 uint8, value u8 / 255): marched from 8-byte cell records, bit-identical to marching the fp32 volume dr_ingest_u8 makes of
the same bytes, and within tolerance of the oracle on those ingested values."""
import numpy as np
import pytest
import torch

from helpers import GRAD_TOL, RGBA_TOL, case_inputs, oracle_backward_views, oracle_forward_views, rel_l2
from oracle import aux_ref

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _raw(shape, seed):
    """A smooth uint8 volume (the synthetic field quantised) with the reference's on-disk axis order [Z][Y][X]."""
    vol, _, _, _ = case_inputs(shape, (8, 8), 8, seed=seed, jitter=False)
    v = (vol[0].numpy() * 255.0 + 0.5).astype(np.uint8)                      # (D,H,W) = [Y][Z][X]
    return np.ascontiguousarray(np.swapaxes(v, 0, 1))                       # on disk: [Z][Y][X]


def _march(vr, vol_lin, tf_r4, cams, jit, sr=1.0, nondiff=False, skip=None):
    b = vr.brick(vol_lin)
    return (b,) + tuple(vr.march(b, tf_r4, cams, sr, jit, nondiff=nondiff, skip=skip))


@pytest.mark.parametrize("shape,tf_name,sr,nondiff", [((40, 36, 44), "tf1", 1.0, False), ((33, 29, 37), "rand", 0.7, False),
                                                     ((48, 48, 48), "tf1", 4.0, True), ((1100, 6, 6), "tf1", 1.0, False),
                                                     ((6, 6, 1100), "rand", 1.0, False)])
def test_u8_march_is_bit_identical_to_the_ingested_fp32_volume(shape, tf_name, sr, nondiff):
    from differender_b200 import VolumeRaycaster
    from differender_b200.utils import volume_from_raw_u8
    out_shape = (40, 32)
    _, tf, cams, jit = case_inputs(shape, out_shape, 64, seed=7, tf_name=tf_name, views=2)
    raw = _raw(shape, 7)
    v32 = volume_from_raw_u8(raw, swap_axes01=True, dtype=torch.float32)          # (1, D, H, W) fp32 = u8 / 255
    v8 = volume_from_raw_u8(raw, swap_axes01=True, dtype=torch.uint8)
    assert v8.dtype == torch.uint8 and v8.shape == v32.shape
    assert np.array_equal(v32[0].cpu().numpy(), aux_ref.ingest_u8(raw, True))
    D, H, W = shape
    tf_r4 = tf.to(DEV).t().contiguous()[None]
    c = cams.to(DEV).contiguous()
    j = None if nondiff else jit.to(DEV).contiguous()
    vr32 = VolumeRaycaster((W, D, H), out_shape, max_samples=4096, tf_resolution=64, layout="cell8")
    vr8 = VolumeRaycaster((W, D, H), out_shape, max_samples=4096, tf_resolution=64)            # auto -> cell8 for uint8
    b32, o32, K32, T32 = _march(vr32, v32.reshape(1, D, H, W), tf_r4, c, j, sr, nondiff)
    b8, o8, K8, T8 = _march(vr8, v8.reshape(1, D, H, W), tf_r4, c, j, sr, nondiff)
    assert b8.dtype == torch.uint8 and b8.shape == (1, D * H * W, 8) and b8.element_size() * 8 == 8       # 8-byte records
    assert torch.equal(o8, o32) and torch.equal(K8, K32) and (nondiff or torch.equal(T8, T32))
    # with and without the skip grid (its min/max come from the converted bytes)
    _, o8n, K8n, _ = _march(vr8, v8.reshape(1, D, H, W), tf_r4, c, j, sr, nondiff, skip=False)
    assert torch.equal(o8, o8n) and torch.equal(K8, K8n)
    if nondiff:
        return
    go = torch.randn(o8.shape, generator=torch.Generator().manual_seed(4)).to(DEV)
    gv32, gt32 = vr32.march_backward(b32, tf_r4, c, sr, j, go, o32, K32, T32, True, True)
    gv8, gt8 = vr8.march_backward(b8, tf_r4, c, sr, j, go, o8, K8, T8, True, True)
    assert gv8.dtype == torch.float32
    assert rel_l2(gv8.cpu().numpy(), gv32.cpu().numpy()) <= 1e-5 and rel_l2(gt8.cpu().numpy(), gt32.cpu().numpy()) <= 1e-5


def test_u8_volume_through_the_public_api_matches_the_oracle():
    from differender_b200 import Raycaster
    from differender_b200.utils import volume_from_raw_u8
    shape, out_shape = (40, 36, 44), (56, 40)
    _, tf, cams, jit = case_inputs(shape, out_shape, 64, seed=9, tf_name="tf1", views=2)
    raw = _raw(shape, 9)
    v8 = volume_from_raw_u8(raw, swap_axes01=True, dtype=torch.uint8)
    vol_f = torch.tensor(aux_ref.ingest_u8(raw, True))[None]                  # what the oracle marches: the ingested fp32 values
    rc = Raycaster(shape, out_shape, 64, max_samples=2048)
    t = tf.to(DEV).requires_grad_(True)
    img = rc(v8, t, cams.to(DEV), jit.to(DEV))                                # a uint8 volume cannot require grad: TF gradient only
    ref, Kr, _ = oracle_forward_views(vol_f, tf, cams, out_shape, jit, max_samples=2048)
    same = rc.vr.last_K.cpu().numpy() == Kr
    assert (~same).mean() <= 1e-4
    assert np.moveaxis(np.abs(img.detach().cpu().numpy() - ref), 1, 0)[:, same].max() <= RGBA_TOL
    go = torch.randn(img.shape, generator=torch.Generator().manual_seed(5))
    (img * go.to(DEV)).sum().backward()
    _, gt = oracle_backward_views(vol_f, tf, cams, go.numpy(), out_shape, jit, max_samples=2048, want_vol=False)
    assert rel_l2(t.grad.cpu().numpy(), gt) <= GRAD_TOL
    nd = rc.raycast_nondiff(v8, tf.to(DEV), cams.to(DEV))
    refn, _, _ = oracle_forward_views(vol_f, tf, cams, out_shape, None, sampling_rate=4.0, nondiff=True)
    assert np.abs(nd.cpu().numpy() - refn).max() <= RGBA_TOL


def test_u8_rejects_other_layouts():
    from differender_b200 import VolumeRaycaster
    vr = VolumeRaycaster((16, 16, 16), (16, 16), tf_resolution=8, layout="linear")
    with pytest.raises(ValueError, match="uint8 volumes"):
        vr.brick(torch.zeros((1, 16, 16, 16), dtype=torch.uint8, device=DEV))
