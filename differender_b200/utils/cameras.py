"""Camera helpers of the reference (differender/utils/utils.py:80-90)."""
import math

import torch
import torch.nn.functional as F

__all__ = ['in_circles', 'get_rand_pos']


def in_circles(i, y=0.7, dist=2.5):
    """Orbit camera (utils.py:80-83)."""
    x = math.cos(i) * dist
    z = math.sin(i) * dist
    return torch.tensor([x, y, z], dtype=torch.float32)


def get_rand_pos(bs=None, dist=2.7):
    """Random camera(s) on a sphere of radius `dist` (utils.py:86-90)."""
    if bs is None:
        return F.normalize(torch.randn(3), dim=0) * dist
    else:
        return F.normalize(torch.randn(bs, 3), dim=1) * dist
