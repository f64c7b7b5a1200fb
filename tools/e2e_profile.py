"""cProfile of the host side of the e2e leg of bench.py (public autograd API with host buffers):
where the CPU time of a step goes.  `python tools/e2e_profile.py [--steps 200]`"""
import argparse
import cProfile
import os
import pstats
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from guassianhand_b200 import api, scenes  # noqa: E402
import util  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=200)
ap.add_argument("--views", type=int, default=8)
a = ap.parse_args()
dev = torch.device("cuda", 0)
P, H, W, B = 60000, 512, 334, a.views
sc = scenes.two_hand_scene(P, seed=0)
cams = scenes.fibonacci_cameras(64, H, W, seed=0)
bg = np.zeros(3, np.float32)
t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).float().to(dev)
gauss = dict(means3D=t(sc.means3D), opacities=t(sc.opacities), scales=t(sc.scales), rotations=t(sc.rotations),
             colors_precomp=t(sc.colors))
views = util.gpu_views(cams[:B], bg, dev)
dL = t((np.random.default_rng(1).normal(size=(B, 3, H, W)) / (H * W)).astype(np.float32))


def step():
    leaf = {k: v.detach().requires_grad_(True) for k, v in gauss.items()}
    imgs, _ = api.rasterize_views(leaf["means3D"], leaf["opacities"], views, colors_precomp=leaf["colors_precomp"],
                                  scales=leaf["scales"], rotations=leaf["rotations"], check="deferred")
    loss = (imgs * dL).sum()
    loss.backward()
    return loss


for _ in range(10):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(a.steps):
    step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host enqueue {1e6*(t1-t0)/a.steps:.1f} us/step, with drain {1e6*(t2-t0)/a.steps:.1f} us/step")
pr = cProfile.Profile()
pr.enable()
for _ in range(a.steps):
    step()
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
