"""Drop-in module name: `from diff_gaussian_rasterization import GaussianRasterizationSettings,
GaussianRasterizer` (/root/reference/tgs/models/renderer_one_shot.py:3) resolves here when this
repository is on sys.path, and runs on libghr.so (sm_100a) instead of the upstream extension."""
from guassianhand_b200.api import (GaussianRasterizationSettings, GaussianRasterizer,  # noqa: F401
                                   rasterize_gaussians)

__all__ = ["GaussianRasterizationSettings", "GaussianRasterizer", "rasterize_gaussians"]
