"""Per-stage CUDA-event durations of the fit step (fwd+bwd) for a given number of views, plus the
whole-step time eager and graphed.  `python tools/stage_times.py --views 1 8 [--P 60000 --H 512 --W 334]`"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from guassianhand_b200 import _native as NV, build as _build, scenes  # noqa: E402
if os.environ.get("GHR_TOOL_VARIANT"):      # A/B of a build variant (guassianhand_b200/build.py VARIANTS); tools only
    NV.LIB_PATH = _build.variant_path(os.environ["GHR_TOOL_VARIANT"])
from guassianhand_b200.dist import GraphedFitStep, PackedGrads, fit_step_grads  # noqa: E402
import util  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--views", type=int, nargs="+", default=[1, 8])
ap.add_argument("--steps", type=int, default=30)
ap.add_argument("--P", type=int, default=60000)
ap.add_argument("--H", type=int, default=512)
ap.add_argument("--W", type=int, default=334)
ap.add_argument("--sh", type=int, default=-1, help="SH degree (-1 = colors_precomp)")
ap.add_argument("--no-flush", action="store_true")
ap.add_argument("--overlap", type=int, nargs="+", default=[1], help="view groups on concurrent streams (graph timing)")
a = ap.parse_args()
dev = torch.device("cuda", 0)
sc = scenes.two_hand_scene(a.P, seed=0, sh_degree=a.sh if a.sh >= 0 else None)
cams = scenes.fibonacci_cameras(64, a.H, a.W, seed=0)
t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).float().to(dev)
gauss = dict(means3D=t(sc.means3D), opacities=t(sc.opacities), scales=t(sc.scales), rotations=t(sc.rotations))
M = 0
if a.sh >= 0:
    gauss["shs"] = t(sc.shs)
    M = sc.shs.shape[1]
else:
    gauss["colors_precomp"] = t(sc.colors)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for V in a.views:
    views = util.gpu_views(cams[:V], np.zeros(3, np.float32), dev)
    dL = t((np.random.default_rng(1).normal(size=(V, 3, a.H, a.W)) / (a.H * a.W)).astype(np.float32))
    grads = PackedGrads(a.P, M, device=dev)
    kw = dict(sh_degree=max(a.sh, 0))
    res = fit_step_grads(gauss, views, dL, grads, **kw)
    cap = int(res.R * 1.25) + (1 << 14)
    for _ in range(3):
        fit_step_grads(gauss, views, dL, grads, R_cap=cap, check="none", **kw)
    fe = [NV.StageEvents(NV.GHR_NSTAGES_FWD) for _ in range(a.steps)]
    be = [NV.StageEvents(NV.GHR_NSTAGES_BWD) for _ in range(a.steps)]
    s0 = [torch.cuda.Event(enable_timing=True) for _ in range(a.steps)]
    s1 = [torch.cuda.Event(enable_timing=True) for _ in range(a.steps)]
    for i in range(a.steps):
        if not a.no_flush:
            flush.zero_()
        s0[i].record()
        fit_step_grads(gauss, views, dL, grads, R_cap=cap, check="none", fwd_events=fe[i], bwd_events=be[i], **kw)
        s1[i].record()
    torch.cuda.synchronize()
    out = {"V": V, "R": res.R, "eager_ms": float(np.mean([x.elapsed_time(y) for x, y in zip(s0, s1)]))}
    for si, name in enumerate(NV.FWD_STAGES):
        out[name] = round(float(np.mean([e.elapsed_ms(si) for e in fe])) * 1000, 1)
    for si, name in enumerate(NV.BWD_STAGES):
        out[name] = round(float(np.mean([e.elapsed_ms(si) for e in be])) * 1000, 1)
    for G in a.overlap:
        G = max(1, min(G, V))
        caps = cap
        if G > 1:
            r = fit_step_grads(gauss, views, dL, grads, overlap=G, **kw)
            caps = [int(x.R * 1.25) + (1 << 14) for x in r.results]
        g = GraphedFitStep(gauss, views, dL, grads, R_cap=caps, overlap=G, **kw)
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        for i in range(a.steps):
            if not a.no_flush:
                flush.zero_()
            s0[i].record()
            g.replay()
            s1[i].record()
        torch.cuda.synchronize()
        if g.status()[1]:
            raise RuntimeError("overflow in the graphed step")
        ms = float(np.mean([x.elapsed_time(y) for x, y in zip(s0, s1)]))
        tag = "" if G == 1 else f"_overlap{G}"
        out["graph_ms" + tag] = ms
        out["views_per_s_graph" + tag] = V / ms * 1000
    print(json.dumps(out))
