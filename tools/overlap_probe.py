"""Experiment: run the V views of a fit step as G independent groups on G streams (captured into one
CUDA graph) so the latency-bound binning chain of one group overlaps the blend kernels of another."""
import argparse, os, sys, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from guassianhand_b200 import scenes
from guassianhand_b200.dist import PackedGrads, fit_step_grads
import util
ap = argparse.ArgumentParser()
ap.add_argument("--views", type=int, default=8)
ap.add_argument("--groups", type=int, nargs="+", default=[1, 2, 4, 8])
ap.add_argument("--steps", type=int, default=50)
ap.add_argument("--prio", type=int, default=0)
a = ap.parse_args()
dev = torch.device("cuda", 0)
P, H, W = 60000, 512, 334
sc = scenes.two_hand_scene(P, seed=0)
cams = scenes.fibonacci_cameras(64, H, W, seed=0)
t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).float().to(dev)
gauss = dict(means3D=t(sc.means3D), opacities=t(sc.opacities), scales=t(sc.scales), rotations=t(sc.rotations), colors_precomp=t(sc.colors))
V = a.views
dL = t((np.random.default_rng(1).normal(size=(V, 3, H, W)) / (H * W)).astype(np.float32))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for G in a.groups:
    per = V // G
    vgs = [util.gpu_views(cams[g * per:(g + 1) * per], np.zeros(3, np.float32), dev) for g in range(G)]
    dLs = [dL[g * per:(g + 1) * per].contiguous() for g in range(G)]
    grads = [PackedGrads(P, 0, device=dev) for _ in range(G)]
    streams = [torch.cuda.Stream(priority=(-1 if a.prio else 0)) for _ in range(G)]
    caps = []
    for g in range(G):
        r = fit_step_grads(gauss, vgs[g], dLs[g], grads[g])
        caps.append(int(r.R * 1.25) + (1 << 14))
    torch.cuda.synchronize()
    def step():
        main = torch.cuda.current_stream()
        for g in range(G):
            streams[g].wait_stream(main)
            with torch.cuda.stream(streams[g]):
                fit_step_grads(gauss, vgs[g], dLs[g], grads[g], R_cap=caps[g], check="none")
        for g in range(G):
            main.wait_stream(streams[g])
        for g in range(1, G):
            grads[0].flat.add_(grads[g].flat)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            step()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    ref = grads[0].flat.clone()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        step()
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize()
    s0 = [torch.cuda.Event(enable_timing=True) for _ in range(a.steps)]
    s1 = [torch.cuda.Event(enable_timing=True) for _ in range(a.steps)]
    for i in range(a.steps):
        flush.zero_()
        s0[i].record(); graph.replay(); s1[i].record()
    torch.cuda.synchronize()
    ms = float(np.mean([x.elapsed_time(y) for x, y in zip(s0, s1)]))
    err = float((grads[0].flat - ref).abs().max() / ref.abs().max())
    print(json.dumps({"V": V, "G": G, "graph_ms": ms, "views_per_s": V / ms * 1000, "replay_vs_eager_relerr": err}))
