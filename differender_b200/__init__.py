"""differender_b200 -- B200-native (sm_100a) implementation of Differender's differentiable ray-march hot path.

Drop-in for `differender.volume_raycaster` (reference differender/__init__.py exposes the same names)."""
from .volume_raycaster import Raycaster, RaycastFunction, RaycastMSEFunction, VolumeRaycaster  # noqa: F401
from .optim import FusedVolumeSGD, MomentumSGD  # noqa: F401
from . import losses  # noqa: F401

__version__ = "0.1.0"
