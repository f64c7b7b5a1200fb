"""Minimal driver for ncu: a few steps of the bench workload (B views of the C2 scene, fwd+bwd)."""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from guassianhand_b200 import scenes  # noqa: E402
from guassianhand_b200.dist import PackedGrads, fit_step_grads  # noqa: E402
import util  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--views", type=int, default=8)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--P", type=int, default=60000)
ap.add_argument("--H", type=int, default=512)
ap.add_argument("--W", type=int, default=334)
a = ap.parse_args()
dev = torch.device("cuda", 0)
sc = scenes.two_hand_scene(a.P, seed=0)
cams = scenes.fibonacci_cameras(64, a.H, a.W, seed=0)
t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).float().to(dev)
gauss = dict(means3D=t(sc.means3D), opacities=t(sc.opacities), scales=t(sc.scales), rotations=t(sc.rotations),
             colors_precomp=t(sc.colors))
views = util.gpu_views(cams[:a.views], np.zeros(3, np.float32), dev)
dL = t((np.random.default_rng(1).normal(size=(a.views, 3, a.H, a.W)) / (a.H * a.W)).astype(np.float32))
grads = PackedGrads(a.P, 0, device=dev)
res = fit_step_grads(gauss, views, dL, grads)
cap = int(res.R * 1.25) + (1 << 14)
for _ in range(a.steps):
    fit_step_grads(gauss, views, dL, grads, R_cap=cap, check="none")
torch.cuda.synchronize()
print("R", res.R)
