"""cProfile of the host side of the drop-in single-view path at the reference-as-shipped shape
(98,562 Gaussians, 256x256, RGB + fused mask, fwd+bwd per call).  `python tools/dropin_profile.py`"""
import cProfile
import os
import pstats
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from guassianhand_b200 import scenes  # noqa: E402
from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer  # noqa: E402

dev = torch.device("cuda", 0)
sc = scenes.two_hand_scene(98562, seed=0)
cam = scenes.fibonacci_cameras(1, 256, 256, seed=0)[0]
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).float().to(dev)
leafs = [t(x).requires_grad_(True) for x in (sc.means3D, sc.opacities, sc.scales, sc.rotations, sc.colors)]
w = t((np.random.default_rng(5).normal(size=(3, 256, 256)) / 65536).astype(np.float32))
view, proj, campos, bg = t(cam.viewmatrix), t(cam.projmatrix), t(cam.campos), torch.zeros(3, device=dev)


ones = torch.ones_like(leafs[0])


def step_two_calls():
    """the reference's pattern unchanged: RGB render, then an all-ones mask render of the same geometry"""
    xyz, op, scl, rot, col = leafs
    rs = GaussianRasterizationSettings(image_height=cam.H, image_width=cam.W, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy,
                                       bg=bg, scale_modifier=1.0, viewmatrix=view, projmatrix=proj, sh_degree=0,
                                       campos=campos, prefiltered=False, debug=False)
    m2d = torch.zeros_like(xyz, requires_grad=True)
    r = GaussianRasterizer(raster_settings=rs)
    img, _ = r(means3D=xyz, means2D=m2d, shs=None, colors_precomp=col, opacities=op, scales=scl, rotations=rot,
               cov3D_precomp=None)
    msk, _ = r(means3D=xyz, means2D=m2d, shs=None, colors_precomp=ones, opacities=op, scales=scl, rotations=rot,
               cov3D_precomp=None)
    loss = (img * w).sum() + (msk[0] * w[0]).sum()
    loss.backward()


def step_fused():
    xyz, op, scl, rot, col = leafs
    rs = GaussianRasterizationSettings(image_height=cam.H, image_width=cam.W, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy,
                                       bg=bg, scale_modifier=1.0, viewmatrix=view, projmatrix=proj, sh_degree=0,
                                       campos=campos, prefiltered=False, debug=False)
    m2d = torch.zeros_like(xyz, requires_grad=True)
    img, _, msk = GaussianRasterizer(raster_settings=rs).forward_with_mask(
        means3D=xyz, means2D=m2d, opacities=op, colors_precomp=col, scales=scl, rotations=rot)
    loss = (img * w).sum() + (msk * w[0]).sum()
    loss.backward()


step = step_two_calls if "--fused" not in sys.argv else step_fused
for _ in range(10):
    step()
torch.cuda.synchronize()
n = 200
t0 = time.perf_counter()
for _ in range(n):
    step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host enqueue {1e6*(t1-t0)/n:.1f} us/step, with drain {1e6*(t2-t0)/n:.1f} us/step")
pr = cProfile.Profile()
pr.enable()
for _ in range(n):
    step()
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(26)
from guassianhand_b200 import api  # noqa: E402
print("geometry cache hits:", api._GeomCache.hits)
# device time of one pair (events), to compare with the host enqueue time
s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
s0.record()
for _ in range(50):
    step()
s1.record()
torch.cuda.synchronize()
print(f"device+host pipeline {s0.elapsed_time(s1) * 1000 / 50:.1f} us/pair")
