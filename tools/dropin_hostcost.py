"""Where the host time of the drop-in call pair goes (reference-as-shipped shape: 98,562 Gaussians, 256x256):
wall-clock per piece with the GPU kept busy-free (each piece is timed over many repetitions, device drained
before and after).  `python tools/dropin_hostcost.py`"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from guassianhand_b200 import api, scenes  # noqa: E402
from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer  # noqa: E402

dev = torch.device("cuda", 0)
sc = scenes.two_hand_scene(98562, seed=0)
cam = scenes.fibonacci_cameras(1, 256, 256, seed=0)[0]
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).float().to(dev)
leafs = [t(x).requires_grad_(True) for x in (sc.means3D, sc.opacities, sc.scales, sc.rotations, sc.colors)]
w = t((np.random.default_rng(5).normal(size=(3, 256, 256)) / 65536).astype(np.float32))
view, proj, campos, bg = t(cam.viewmatrix), t(cam.projmatrix), t(cam.campos), torch.zeros(3, device=dev)
ones = torch.ones_like(leafs[0])
rs = GaussianRasterizationSettings(image_height=cam.H, image_width=cam.W, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy,
                                   bg=bg, scale_modifier=1.0, viewmatrix=view, projmatrix=proj, sh_degree=0,
                                   campos=campos, prefiltered=False, debug=False)
xyz, op, scl, rot, col = leafs
N = 300


def timed(name, fn, n=N, batch=12):
    """host: enqueue time in batches short enough that the launch queue (1024 entries) never fills -- a full
    queue makes every launch wait for the GPU and the host time reads as device time; device: one long run."""
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    host = 0.0
    for _ in range(n // batch):
        t0 = time.perf_counter()
        for _ in range(batch):
            fn()
        host += time.perf_counter() - t0
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"{name:58s} host {1e6 * host / (n // batch * batch):7.1f} us   pipelined {1e6 * (t2 - t0) / n:7.1f} us", flush=True)


class _Ident(torch.autograd.Function):
    """autograd node with the rasterizer's signature and NO work: what autograd + the harness cost by themselves"""

    @staticmethod
    def forward(ctx, means3D, means2D, colors, opac, scales, rots):
        ctx.save_for_backward(means3D, opac, scales, rots, colors)
        return w * 1.0, torch.zeros(1, device=dev)

    @staticmethod
    def backward(ctx, g, _):
        m, o, s, r, c = ctx.saved_tensors
        return torch.zeros_like(m), torch.zeros_like(m), torch.zeros_like(c), torch.zeros_like(o), torch.zeros_like(s), torch.zeros_like(r)


def harness_only():
    m2d = torch.zeros_like(xyz, requires_grad=True)
    img, _ = _Ident.apply(xyz, m2d, col, op, scl, rot)
    msk, _ = _Ident.apply(xyz, m2d, ones, op, scl, rot)
    loss = (img * w).sum() + (msk[0] * w[0]).sum()
    loss.backward()


def pair():
    m2d = torch.zeros_like(xyz, requires_grad=True)
    r = GaussianRasterizer(raster_settings=rs)
    img, _ = r(means3D=xyz, means2D=m2d, shs=None, colors_precomp=col, opacities=op, scales=scl, rotations=rot,
               cov3D_precomp=None)
    msk, _ = r(means3D=xyz, means2D=m2d, shs=None, colors_precomp=ones, opacities=op, scales=scl, rotations=rot,
               cov3D_precomp=None)
    loss = (img * w).sum() + (msk[0] * w[0]).sum()
    loss.backward()


def fwd_only_nograd():
    with torch.no_grad():
        r = GaussianRasterizer(raster_settings=rs)
        r(means3D=xyz, means2D=xyz, shs=None, colors_precomp=col, opacities=op, scales=scl, rotations=rot,
          cov3D_precomp=None)
        r(means3D=xyz, means2D=xyz, shs=None, colors_precomp=ones, opacities=op, scales=scl, rotations=rot,
          cov3D_precomp=None)


def fwd_only_grad():
    m2d = torch.zeros_like(xyz, requires_grad=True)
    r = GaussianRasterizer(raster_settings=rs)
    r(means3D=xyz, means2D=m2d, shs=None, colors_precomp=col, opacities=op, scales=scl, rotations=rot,
      cov3D_precomp=None)
    r(means3D=xyz, means2D=m2d, shs=None, colors_precomp=ones, opacities=op, scales=scl, rotations=rot,
      cov3D_precomp=None)


cams = api._cams_from_settings(rs)
det = [x.detach() for x in leafs]


def raw_fwd_pair():
    a = api.forward_raw(cams, det[0], det[1], det[2], det[3], None, None, det[4], 0, 1.0, check="auto")
    api.forward_raw(cams, det[0], det[1], det[2], det[3], None, None, ones, 0, 1.0, R_cap=a.R_cap,
                    reuse=(a.state, 0, a.R))
    return a


state = raw_fwd_pair()
dL = w.unsqueeze(0).contiguous()


def raw_bwd_pair():
    for c in (det[4], ones):
        api.backward_raw(cams, state.state, state.R_cap, dL, det[0], det[1], det[2], det[3], None, None, c, 0, 1.0)


def raw_fwd_bwd_pair():
    a = raw_fwd_pair()
    for c in (det[4], ones):
        api.backward_raw(cams, a.state, a.R_cap, dL, det[0], det[1], det[2], det[3], None, None, c, 0, 1.0)


timed("harness only (identity autograd nodes, loss, backward)", harness_only)
timed("raw forward x2 (forward_raw + reuse; no autograd)", raw_fwd_pair)
timed("raw backward x2 (backward_raw; no autograd)", raw_bwd_pair)
timed("raw forward x2 + backward x2 (device-bound floor)", raw_fwd_bwd_pair)
timed("drop-in forward x2, no_grad", fwd_only_nograd)
timed("drop-in forward x2, autograd recording", fwd_only_grad)
timed("drop-in pair fwd + bwd (the as_shipped two_calls step)", pair)

# ---- inside the raw calls: time spent in the C entry points themselves (kernel launches) vs Python around them
L = api.N.lib()
acc = {"ghr_forward": [0.0, 0], "ghr_backward": [0.0, 0]}


def wrap(name):
    fn = getattr(L, name)

    def timed_fn(*a):
        t0 = time.perf_counter()
        r = fn(*a)
        acc[name][0] += time.perf_counter() - t0
        acc[name][1] += 1
        return r
    return timed_fn


class _Proxy:
    def __init__(self, lib_):
        self._l = lib_
        self.ghr_forward = wrap("ghr_forward")
        self.ghr_backward = wrap("ghr_backward")

    def __getattr__(self, k):
        return getattr(self._l, k)


api.N._lib = _Proxy(L)
torch.cuda.synchronize()
t1 = t0 = 0.0
for _ in range(N // 12):
    ta = time.perf_counter()
    for _ in range(12):
        raw_fwd_bwd_pair()
    t1 += time.perf_counter() - ta
    torch.cuda.synchronize()
acc = {k: [v[0] * N / (N // 12 * 12), v[1] * N // (N // 12 * 12)] for k, v in acc.items()}
t1 = t1 * N / (N // 12 * 12)
print(f"raw fwd x2 + bwd x2: host {1e6 * (t1 - t0) / N:.1f} us per pair; inside ghr_forward "
      f"{1e6 * acc['ghr_forward'][0] / N:.1f} us ({acc['ghr_forward'][1] // N} calls), inside ghr_backward "
      f"{1e6 * acc['ghr_backward'][0] / N:.1f} us ({acc['ghr_backward'][1] // N} calls)")
api.N._lib = L
import cProfile
import pstats
pr = cProfile.Profile()
pr.enable()
for _ in range(N // 12):
    for _ in range(12):
        raw_fwd_pair()
    pr.disable()
    torch.cuda.synchronize()
    pr.enable()
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
