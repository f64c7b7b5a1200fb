"""Forward-only render throughput (BASELINE config 5: novel-pose drive render, 1920x1080, one pose per
call): eager calls and a CUDA-graph replay.  `python tools/fwd_only.py [--P 60000 --H 1080 --W 1920]`"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from guassianhand_b200 import api, scenes  # noqa: E402
import util  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--P", type=int, default=60000)
ap.add_argument("--H", type=int, default=1080)
ap.add_argument("--W", type=int, default=1920)
ap.add_argument("--poses", type=int, default=32)
a = ap.parse_args()
dev = torch.device("cuda", 0)
sc = scenes.two_hand_scene(a.P, seed=0)
cams = scenes.fibonacci_cameras(a.poses, a.H, a.W, seed=0)
t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).float().to(dev)
rng = np.random.default_rng(0)
# one Gaussian set per pose: rigid jitter + noise of the means
means = [t(sc.means3D + rng.normal(0, 0.002, size=(1, 3)).astype(np.float32)
           + rng.normal(0, 0.0005, size=sc.means3D.shape).astype(np.float32)) for _ in range(a.poses)]
opac, scl, rot, col = t(sc.opacities), t(sc.scales), t(sc.rotations), t(sc.colors)
views = [util.gpu_views([c], np.zeros(3, np.float32), dev).cams() for c in cams]


def render(i, cap=None, check="poll"):
    return api.forward_raw(views[i], means[i], opac, scl, rot, None, None, col, 0, 1.0, R_cap=cap, check=check)


Rs = [render(i).R for i in range(a.poses)]
cap = int(max(Rs) * 1.25) + (1 << 14)
for i in range(a.poses):
    render(i, cap, "none")
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for i in range(a.poses):
    render(i, cap, "none")
e.record()
torch.cuda.synchronize()
eager_ms = s.elapsed_time(e) / a.poses
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    outs = [render(i, cap, "none") for i in range(a.poses)]
for _ in range(2):
    g.replay()
torch.cuda.synchronize()
s.record()
for _ in range(5):
    g.replay()
e.record()
torch.cuda.synchronize()
graph_ms = s.elapsed_time(e) / (5 * a.poses)
print(json.dumps({"P": a.P, "H": a.H, "W": a.W, "poses": a.poses, "R_mean": float(np.mean(Rs)),
                  "eager_ms_per_pose": eager_ms, "poses_per_s_eager": 1000 / eager_ms,
                  "graph_ms_per_pose": graph_ms, "poses_per_s_graph": 1000 / graph_ms}))
