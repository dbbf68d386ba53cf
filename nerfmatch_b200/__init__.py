"""nerfmatch_b200 - B200-native scene-feature renderer for NeRFMatch (drop-in for nerfmatch/nerf)."""
from ._lib import NmrError, build, load  # noqa: F401
from .renderer import NerfRenderer  # noqa: F401

__all__ = ["NerfRenderer", "NmrError", "build", "load"]
