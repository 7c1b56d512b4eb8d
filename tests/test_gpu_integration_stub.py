"""The ctypes stub printed in INTEGRATION.md (what a maintainer of the reference would add) must actually work: it is
extracted from the document, pointed at the in-tree library, and its forward/backward are compared with the product host."""
import os
import re
import types

import numpy as np
import pytest
import torch

from helpers import case_inputs, rel_l2

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_integration_stub_runs_and_matches():
    from differender_b200 import Raycaster, _lib
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    code = re.search(r"```python\n(# differender/_diffrender\.py.*?)```", doc, flags=re.S).group(1)
    code = code.replace('ctypes.CDLL("libdiffrender.so")', f'ctypes.CDLL({_lib.LIB_PATH!r})')
    mod = types.ModuleType("stub")
    exec(compile(code, "INTEGRATION.md", "exec"), mod.__dict__)
    out_shape = (40, 24)
    vol, tf, cams, _ = case_inputs((32, 28, 36), out_shape, 32, seed=41, tf_name="tf1", views=1, jitter=False)
    rc = Raycaster((32, 28, 36), out_shape, 32, jitter=False, max_samples=1024)
    dev = "cuda:0"
    v = vol.to(dev).requires_grad_(True); t = tf.to(dev).requires_grad_(True)
    _, _, vol_in, tf_in, lf_in = rc._determine_batch(v, t, cams[0].to(dev))
    raw = mod.RaycastFunction.apply(rc.vr, vol_in, tf_in, lf_in, 1.0, (False, 0), False)        # reference calling convention
    assert raw.shape == (40, 24, 4)
    img = torch.flip(raw, (1,)).permute(2, 1, 0).contiguous()                                    # reference :543-548
    go = torch.randn(img.shape, generator=torch.Generator().manual_seed(2)).to(dev)
    (img * go).sum().backward()
    v2 = vol.to(dev).requires_grad_(True); t2 = tf.to(dev).requires_grad_(True)
    img2 = rc(v2, t2, cams[0].to(dev))
    (img2 * go).sum().backward()
    assert torch.equal(img.detach(), img2.detach())
    assert rel_l2(v.grad.cpu().numpy(), v2.grad.cpu().numpy()) <= 1e-5
    assert rel_l2(t.grad.cpu().numpy(), t2.grad.cpu().numpy()) <= 1e-5
