"""Host-side mirror of the reference interface: shapes, batching rule, helpers, error behaviour (no GPU needed)."""
import math

import pytest
import torch

from differender_b200 import Raycaster, RaycastFunction, VolumeRaycaster
from differender_b200.distributed import shard_views
from differender_b200.utils import get_rand_pos, get_tf, in_circles, tex_from_pts


def _rc(**kw):
    return Raycaster((12, 16, 20), (32, 24), 8, **kw)


def test_constructor_mirrors_reference_attributes():
    rc = _rc(sampling_rate=2.0, jitter=False, max_samples=99, fov=45.0, ti_kwargs={"debug": True})
    assert rc.volume_shape == (20, 12, 16)          # torch (D,H,W) -> Taichi (W,D,H)  (reference :481)
    assert rc.output_shape == (32, 24) and rc.tf_shape == 8 and rc.sampling_rate == 2.0 and rc.jitter is False
    assert isinstance(rc.vr, VolumeRaycaster) and rc.vr.max_samples == 99 and rc.vr.resolution == (32, 24)
    assert "Max Samples = 99" in repr(rc) and "Volume ((20, 12, 16))" in repr(rc)


def test_determine_batch_rule_and_views():
    rc = _rc()
    vol, tf, lf = torch.rand(1, 12, 16, 20), torch.rand(4, 8), torch.rand(3)
    b, bs, v, t, l = rc._determine_batch(vol, tf, lf)
    assert (b, bs) == (False, 0) and v.shape == (20, 12, 16) and t.shape == (8, 4) and l.shape == (3,)
    assert v.data_ptr() == vol.data_ptr()                          # a view, not a copy
    assert torch.equal(v[3, 5, 7], vol[0, 5, 7, 3])                 # Taichi [x,y,z] == torch [0, d=y, h=z, w=x]
    b, bs, v, t, l = rc._determine_batch(vol, tf, torch.rand(5, 3))
    assert (b, bs) == (True, 5) and v.shape == (20, 12, 16) and t.shape == (8, 4)      # shared, NOT cloned 5x (:566)
    b, bs, v, t, l = rc._determine_batch(torch.rand(3, 1, 12, 16, 20), torch.rand(3, 4, 8), torch.rand(3, 3))
    assert (b, bs) == (True, 3) and v.shape == (3, 20, 12, 16) and t.shape == (3, 8, 4)
    b, bs, v, t, l = rc._determine_batch(vol, torch.rand(2, 4, 8), lf)
    assert (b, bs) == (True, 2)                                     # batch size from the first batched input


def test_cpu_tensors_raise_no_fallback():
    rc = _rc()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        rc(torch.rand(1, 12, 16, 20), torch.rand(4, 8), torch.rand(3))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        rc.raycast_nondiff(torch.rand(1, 12, 16, 20), torch.rand(4, 8), torch.rand(3))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        RaycastFunction.apply(rc.vr, torch.rand(20, 12, 16), torch.rand(8, 4), torch.rand(3), 1.0, (False, 0), False)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from differender_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU or PyTorch fallback"):
        _lib.load()


def test_product_never_imports_the_oracle():
    import os, re
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "differender_b200")
    for dp, _, files in os.walk(root):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "hostsim" not in src or f == "dr_math.cuh" or f == "dr_desc.h", f


def test_helpers_match_reference_presets():
    t = get_tf("tf1", 128)
    assert t.shape == (4, 128) and t.dtype == torch.float32
    assert torch.all(t[:, 0] == 0) and torch.all(t[:, -1] == 0)                  # first and last control points are zero
    assert abs(float(t[3].max()) - 0.3917) < 2e-3                                  # plateau of the last alpha bump
    pts = torch.tensor([[0.0, 0, 0, 0, 0], [0.5, 1, 1, 1, 1], [1.0, 0, 0, 0, 0]])
    tri = tex_from_pts(pts, 5)
    assert torch.allclose(tri[3], torch.tensor([0.0, 0.5, 1.0, 0.5, 0.0]))
    assert torch.all(get_tf("black", 16) == 1e-2) and get_tf("gray", 4)[3, 0] == 0.02 and get_tf("rand", 9).shape == (4, 9)
    with pytest.raises(Exception, match="Invalid Transfer function"):
        get_tf("nope", 4)
    c = in_circles(math.pi / 2)
    assert torch.allclose(c, torch.tensor([0.0, 0.7, 2.5]), atol=1e-6)
    p = get_rand_pos(7)
    assert p.shape == (7, 3) and torch.allclose(p.norm(dim=1), torch.full((7,), 2.7), atol=1e-5)
    assert get_rand_pos().shape == (3,)


def test_shard_views_partitions_exactly():
    for n in (1, 7, 16, 64, 256):
        for world in (1, 2, 3, 8):
            parts = [shard_views(n, r, world) for r in range(world)]
            assert sorted(sum(parts, [])) == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def test_auto_layout_policy():
    # layout="auto": cell-major copy while it stays below AUTO_CELL_BYTES (8x the volume), the 8^3-bricked copy above,
    # the linear tensor in place when an axis is long enough for the generic taps (> 2000 voxels)
    def pick(shape_xyz, dtype, layout="auto", bvol=1):
        X, Y, Z = shape_xyz
        vr = VolumeRaycaster(shape_xyz, (64, 64), layout=layout)
        return vr.resolve_layout(torch.empty(bvol, Y, Z, X, dtype=dtype, device="meta"))
    assert pick((256, 256, 256), torch.float32) == "cell8"                 # C1-C3: 512 MiB
    assert pick((512, 512, 512), torch.float32) == "cell8"                 # C4: 4 GiB
    assert pick((1024, 1024, 1024), torch.float16) == "cell8"              # C5: 16 GiB
    assert pick((1024, 1024, 1024), torch.float32) == "cell8"              # 32 GiB
    assert pick((1536, 1536, 1536), torch.float16) == "brick8"             # 54 GiB of cell-major data: over the limit
    assert pick((512, 512, 512), torch.float32, bvol=16) == "brick8"       # 16 batched volumes: 64 GiB
    assert pick((2100, 64, 64), torch.float32) == "linear"                 # generic taps: linear layout only
    for forced in ("linear", "brick8", "cell8"):
        assert pick((256, 256, 256), torch.float32, layout=forced) == forced
    with pytest.raises(ValueError):
        VolumeRaycaster((8, 8, 8), (8, 8), layout="morton")


def test_uint8_voxel_value_is_exactly_the_fp32_quotient():
    # DR_VOX_U8: the kernels form u8 / 255 as fma(x, r, x * r_lo) (two FMA-pipe instructions, no division); it must round
    # exactly like numpy's float32 division (= dr_ingest_u8, reference examples/taichi_volume_raycaster.py:548-550) for all 256 values
    import ctypes
    import numpy as np
    import hostsim_lib as hs
    out = np.zeros(256, np.float32)
    hs.lib().sim_u8_values(out.ctypes.data_as(ctypes.POINTER(ctypes.c_float)))
    ref = (np.arange(256, dtype=np.float32) / np.float32(255.0)).astype(np.float32)
    assert np.array_equal(out, ref)


def test_helpers_equal_the_reference_utils_where_it_is_mounted():
    """differender/utils/utils.py of the reference, executed as it is with `torchvtk` (un-vendored, not installed) stubbed by this
    package's tex_from_pts: the preset control points, `in_circles` and `get_rand_pos` must be the reference's, value for value."""
    import os, sys, types
    path = os.path.join(os.environ.get("DIFFERENDER_REFERENCE", "/root/reference"), "differender", "utils", "utils.py")
    if not os.path.isfile(path):
        pytest.skip("reference source not mounted (GPU box)")
    stub, stub_utils = types.ModuleType("torchvtk"), types.ModuleType("torchvtk.utils")
    stub_utils.tex_from_pts, stub_utils.TFGenerator = tex_from_pts, None
    stub.utils = stub_utils
    saved = {k: sys.modules.get(k) for k in ("torchvtk", "torchvtk.utils")}
    sys.modules["torchvtk"], sys.modules["torchvtk.utils"] = stub, stub_utils
    try:
        ref = types.ModuleType("differender_reference_utils")
        exec(compile(open(path).read(), path, "exec"), ref.__dict__)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    for name in ("tf1", "tf2", "tf3", "tf4", "tf5", "black", "gray"):
        for res in (16, 128, 257):
            assert torch.equal(ref.get_tf(name, res), get_tf(name, res)), (name, res)
    torch.manual_seed(5); a = ref.get_tf("rand", 64)
    torch.manual_seed(5); b = get_tf("rand", 64)
    assert torch.equal(a, b)
    with pytest.raises(Exception):
        ref.get_tf("nope", 8)
    with pytest.raises(Exception):
        get_tf("nope", 8)
    for i in (0.0, 0.3, math.pi / 2, 4.0):
        assert torch.equal(ref.in_circles(i), in_circles(i)) and torch.equal(ref.in_circles(i, y=0.2, dist=3.0), in_circles(i, y=0.2, dist=3.0))
    for bs in (None, 1, 7):
        torch.manual_seed(9); a = ref.get_rand_pos(bs)
        torch.manual_seed(9); b = get_rand_pos(bs)
        assert torch.equal(a, b)
        torch.manual_seed(9); a = ref.get_rand_pos(bs, dist=1.5)
        torch.manual_seed(9); b = get_rand_pos(bs, dist=1.5)
        assert torch.equal(a, b)
