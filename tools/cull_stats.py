"""Cull efficiency of the blend kernels on the bench workload (counting variant libghr_count.so):
(pixel, instance) pairs evaluated after culling vs pairs that contribute, vs upstream's loop count
I = sum n_contrib.  Prints one JSON object.  usage: python tools/cull_stats.py [--views 8]"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from guassianhand_b200 import _native as NV, build  # noqa: E402
NV.LIB_PATH = build.variant_path(os.environ.get("GHR_COUNT_VARIANT", "count"))   # (A/B: another counting variant)
from guassianhand_b200 import scenes  # noqa: E402
from guassianhand_b200.dist import PackedGrads, fit_step_grads  # noqa: E402
import util  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--views", type=int, default=8)
ap.add_argument("--P", type=int, default=60000)
ap.add_argument("--H", type=int, default=512)
ap.add_argument("--W", type=int, default=334)
a = ap.parse_args()
dev = torch.device("cuda", 0)
L = NV.lib()
sc = scenes.two_hand_scene(a.P, seed=0)
cams = scenes.fibonacci_cameras(64, a.H, a.W, seed=0)
t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).float().to(dev)
gauss = dict(means3D=t(sc.means3D), opacities=t(sc.opacities), scales=t(sc.scales), rotations=t(sc.rotations),
             colors_precomp=t(sc.colors))
V = a.views
views = util.gpu_views(cams[:V], np.zeros(3, np.float32), dev)
dL = t((np.random.default_rng(1).normal(size=(V, 3, a.H, a.W)) / (a.H * a.W)).astype(np.float32))
grads = PackedGrads(a.P, 0, device=dev)
res = fit_step_grads(gauss, views, dL, grads)
cap = int(res.R * 1.25) + (1 << 14)
out = (C.c_ulonglong * 4)()
L.ghr_debug_counts(out, 1)
res = fit_step_grads(gauss, views, dL, grads, R_cap=cap, check="none")
L.ghr_debug_counts(out, 1)
lay = NV.layout(a.P, V, a.H, a.W, 0, 0, cap)
N = a.H * a.W
nc = res.state[lay.off_ncontrib: lay.off_ncontrib + V * N * 4].view(torch.int32)
I = int(nc.sum(dtype=torch.int64).item())
# counters are warp sums: the per-warp survivor count arrives multiplied by 32 lanes already (forward: 32 pixels
# per warp; backward: 64 pixels per warp = x 2)
fe, fc, be, bc = int(out[0]), int(out[1]), int(out[2]) * 2, int(out[3])
print(json.dumps({
    "views": V, "upstream_loop_pairs_I": I, "instances_R": int(res.state[:8].view(torch.int64)[0].item()),
    "forward": {"evaluated_pairs": fe, "contributing_pairs": fc, "evaluated_per_contributing": fe / max(fc, 1), "evaluated_over_I": fe / max(I, 1)},
    "backward": {"evaluated_pairs": be, "contributing_pairs": bc, "evaluated_per_contributing": be / max(bc, 1), "evaluated_over_I": be / max(I, 1)},
}))
