"""Input helpers of the reference (differender/utils/utils.py:7-90) without the torchvtk dependency.

`tex_from_pts` restates torchvtk.utils.tex_from_pts (un-vendored, not installed here) as what its call sites
need: piecewise-linear interpolation of (x, r, g, b, a) control points at `resolution` equidistant positions in
[0,1], returned as a (4, resolution) tensor.  The 'generate' preset (torchvtk.TFGenerator, utils.py:74-77) is not
provided.
"""
import torch

__all__ = ['get_tf', 'tex_from_pts']


def tex_from_pts(tf_pts, resolution):
    """(N, 5) control points [x, r, g, b, a] sorted by x -> (4, resolution) texture."""
    pts = torch.as_tensor(tf_pts, dtype=torch.float64)
    xs, vals = pts[:, 0].contiguous(), pts[:, 1:]
    q = torch.linspace(0.0, 1.0, resolution, dtype=torch.float64)
    hi = torch.searchsorted(xs, q, right=True).clamp(1, len(xs) - 1)
    lo = hi - 1
    span = (xs[hi] - xs[lo])
    w = torch.where(span > 0, (q - xs[lo]) / span.clamp_min(1e-300), torch.zeros_like(q)).clamp(0.0, 1.0)
    tex = vals[lo] * (1.0 - w).unsqueeze(1) + vals[hi] * w.unsqueeze(1)
    return tex.t().contiguous().float()


_PRESETS = {   # control points of utils.py:9-65
    'tf1': [[0.0000, 0.0000, 0.0000, 0.0000, 0.0000], [0.0840, 0.8510, 0.7230, 0.4672, 0.0000],
            [0.0850, 0.8510, 0.7230, 0.4672, 0.0831], [0.1844, 0.8510, 0.7230, 0.4672, 0.0801],
            [0.1890, 0.8510, 0.7230, 0.4672, 0.0000], [0.2444, 0.8667, 0.5166, 0.6566, 0.0000],
            [0.2528, 0.7176, 0.0675, 0.3276, 0.0782], [0.2621, 0.8667, 0.5166, 0.6566, 0.0000],
            [0.3407, 0.9843, 0.9843, 0.9843, 0.0000], [0.3601, 0.9843, 0.9843, 0.9843, 0.3904],
            [0.4475, 0.9843, 0.9843, 0.9843, 0.3917], [0.4655, 0.9843, 0.9843, 0.9843, 0.0000],
            [1.0000, 0.0000, 0.0000, 0.0000, 0.0000]],
    'tf2': [[0.0000, 0.0000, 0.0000, 0.0000, 0.0000], [0.0178, 0.5333, 0.3597, 0.1861, 0.0000],
            [0.0206, 0.5333, 0.3597, 0.1861, 0.1834], [0.0361, 0.5333, 0.3597, 0.1861, 0.1804],
            [0.0388, 0.5333, 0.3597, 0.1861, 0.0000], [0.2224, 0.6902, 0.0839, 0.1951, 0.0000],
            [0.2274, 0.6902, 0.0839, 0.1951, 0.0880], [0.2479, 0.6902, 0.0839, 0.1951, 0.0831],
            [0.2515, 0.6902, 0.0839, 0.1951, 0.0000], [0.2857, 0.9843, 0.9843, 0.9843, 0.0000],
            [0.3042, 0.9843, 0.9843, 0.9843, 0.8240], [0.4540, 0.9843, 0.9843, 0.9843, 0.8172],
            [0.4916, 0.9843, 0.9843, 0.9843, 0.0000], [1.0000, 0.0000, 0.0000, 0.0000, 0.0000]],
    'tf3': [[0.0000, 0.0000, 0.0000, 0.0000, 0.0000], [0.0279, 0.5991, 0.6235, 0.1345, 0.0000],
            [0.0477, 0.5991, 0.6235, 0.1345, 0.1736], [0.1090, 0.5991, 0.6235, 0.1345, 0.1779],
            [0.1304, 0.5991, 0.6235, 0.1345, 0.0000], [0.3654, 0.9843, 0.9843, 0.9843, 0.0000],
            [0.3991, 0.9843, 0.9843, 0.9843, 0.3912], [0.7440, 0.9843, 0.9843, 0.9843, 0.3893],
            [0.7850, 0.9843, 0.9843, 0.9843, 0.0000], [1.0000, 0.0000, 0.0000, 0.0000, 0.0000]],
    'tf4': [[0.0000, 0.0000, 0.0000, 0.0000, 0.0000], [0.0916, 0.5059, 0.1627, 0.1627, 0.0000],
            [0.1204, 0.5059, 0.1627, 0.1627, 0.1932], [0.1865, 0.5059, 0.1627, 0.1627, 0.1956],
            [0.2120, 0.5059, 0.1627, 0.1627, 0.0000], [0.4841, 0.9176, 0.9176, 0.9176, 0.0000],
            [0.5195, 0.9176, 0.9176, 0.9176, 0.6406], [0.6609, 0.9176, 0.9176, 0.9176, 0.6362],
            [0.6968, 0.9176, 0.9176, 0.9176, 0.0000], [1.0000, 0.0000, 0.0000, 0.0000, 0.0000]],
    'tf5': [[0.0000, 0.0000, 0.0000, 0.0000, 0.0000], [0.1300, 0.5000, 0.5000, 0.5000, 0.0000],
            [0.1350, 0.5000, 0.5000, 0.5000, 0.7500], [0.1600, 0.5000, 0.5000, 0.5000, 0.7500],
            [0.1700, 0.5000, 0.5000, 0.5000, 0.0000], [1.0000, 0.0000, 0.0000, 0.0000, 0.0000]],
}


def get_tf(id, res):
    """Preset transfer functions as (4, res) tensors (utils.py:7-79)."""
    if id in _PRESETS:
        return tex_from_pts(torch.tensor(_PRESETS[id]), res)          # fp32 control points, as the reference passes them (utils.py:9)
    elif id == 'black':
        return torch.zeros((4, res)) + 1e-2
    elif id == 'gray':
        temp = torch.ones((4, res)) * 0.5
        temp[3, :] = 0.02
        return temp
    elif id == 'rand':
        return torch.rand(4, res)
    elif id == 'generate':
        raise NotImplementedError("'generate' needs torchvtk.TFGenerator, which is outside the hot path (SURVEY 8(f))")
    else:
        raise Exception(f'Invalid Transfer function identifier given ({id}).')
