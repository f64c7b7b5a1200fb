"""guassianhand_b200 -- B200-native (sm_100a) differentiable Gaussian-splatting rasterizer.

Scope: exactly the hot path GuassianHand delegates to `diff_gaussian_rasterization`
(/root/reference/tgs/models/renderer_one_shot.py:3, :281-346, :355-379).  See DESIGN.md.
"""
from .api import (GaussianRasterizationSettings, GaussianRasterizer, ViewBatch, rasterize_gaussians,
                  rasterize_views, check_deferred)

__all__ = ["GaussianRasterizationSettings", "GaussianRasterizer", "ViewBatch", "rasterize_gaussians",
           "rasterize_views", "check_deferred"]
