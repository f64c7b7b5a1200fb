"""Shared helpers for the parity tests: run the same seeded scene through the CUDA path (via the
C ABI) and through the CPU oracle, and expose every intermediate the contract names."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def oracle_scene(scene, cam, bg, scale_modifier=1.0, cov3D=None):
    from oracle import oracle_lib as ol
    return ol.OracleScene(
        H=cam.H, W=cam.W, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy, bg=bg, viewmatrix=cam.viewmatrix,
        projmatrix=cam.projmatrix, campos=cam.campos, means3D=scene.means3D, opacities=scene.opacities,
        scales=None if cov3D is not None else scene.scales,
        rotations=None if cov3D is not None else scene.rotations, cov3D_precomp=cov3D, shs=scene.shs,
        colors_precomp=scene.colors, sh_degree=scene.sh_degree, scale_modifier=scale_modifier)


def run_oracle(scene, cam, bg, dL=None, scale_modifier=1.0, cov3D=None):
    from oracle import oracle_lib as ol
    osc = oracle_scene(scene, cam, bg, scale_modifier, cov3D)
    if dL is None:
        return ol.forward(osc), None
    return ol.forward_backward(osc, dL)


def gpu_views(cams, bg, device, sh_degree=0, scale_modifier=1.0):
    import torch
    from guassianhand_b200.api import ViewBatch
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).float().to(device)
    return ViewBatch(
        image_height=cams[0].H, image_width=cams[0].W,
        viewmatrix=t(np.stack([c.viewmatrix for c in cams])), projmatrix=t(np.stack([c.projmatrix for c in cams])),
        campos=t(np.stack([c.campos for c in cams])),
        tanfov=t(np.array([[c.tanfovx, c.tanfovy] for c in cams], np.float32)), bg=t(bg),
        sh_degree=sh_degree, scale_modifier=scale_modifier)


def run_gpu(scene, cams, bg, dL=None, scale_modifier=1.0, cov3D=None, R_cap=None, device="cuda:0"):
    """Returns per-view list of dicts (same keys as the oracle's forward dict) and grads dict."""
    import torch
    from guassianhand_b200 import api
    t = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a)).float().to(device)
    views = gpu_views(cams, bg, device, scene.sh_degree, scale_modifier)
    c = views.cams()
    means3D, opac = t(scene.means3D), t(scene.opacities)
    sc = None if cov3D is not None else t(scene.scales)
    rot = None if cov3D is not None else t(scene.rotations)
    cov = t(cov3D)
    shs, col = t(scene.shs), t(scene.colors)
    res = api.forward_raw(c, means3D, opac, sc, rot, cov, shs, col, scene.sh_degree, scale_modifier,
                          want_debug=True, R_cap=R_cap)
    torch.cuda.synchronize()
    lay = res.debug["layout"]
    P, V, H, W = scene.P, len(cams), cams[0].H, cams[0].W
    gx, gy = (W + 15) // 16, (H + 15) // 16
    T = gx * gy
    st = res.state.cpu().numpy()
    geom = st[lay.off_geom: lay.off_geom + V * P * 64].view(np.float32).reshape(V, P, 16)
    ranges = st[lay.off_ranges: lay.off_ranges + V * T * 8].view(np.uint32).reshape(V, T, 2)
    final_T = st[lay.off_final_T: lay.off_final_T + V * H * W * 4].view(np.float32).reshape(V, H, W)
    ncontrib = st[lay.off_ncontrib: lay.off_ncontrib + V * H * W * 4].view(np.uint32).reshape(V, H, W)
    status = st[:32].view(np.uint64)
    R = int(status[0])
    keys = res.debug["keys"].cpu().numpy().view(np.uint64)[:R]
    plist = res.debug["point_list"].cpu().numpy().view(np.uint32)[:R]
    color = res.color.cpu().numpy()
    radii = res.radii.cpu().numpy()
    out = []
    for v in range(V):
        s0 = int(ranges[v][ranges[v, :, 1] > 0, 0].min()) if (ranges[v, :, 1] > 0).any() else 0
        e0 = int(ranges[v, :, 1].max()) if (ranges[v, :, 1] > 0).any() else 0
        rv = ranges[v].astype(np.int64).copy()
        nz = rv[:, 1] > 0
        rv[nz] -= s0                       # per-view offsets, as a single-view upstream call has
        out.append(dict(
            depths=geom[v, :, 11].copy(), radii=radii[v], xy=geom[v, :, 0:2].copy(),
            conic_opacity=geom[v][:, [2, 3, 4, 5]].copy(), rgb=geom[v][:, [8, 9, 10]].copy(),
            tiles_touched=geom[v, :, 13].copy().view(np.uint32), keys=keys[s0:e0], point_list=plist[s0:e0],
            ranges=rv.astype(np.uint32), out_color=color[v], final_T=final_T[v], n_contrib=ncontrib[v], R=e0 - s0))
    grads = None
    if dL is not None:
        g = api.backward_raw(c, res.state, res.R_cap, t(np.asarray(dL).reshape(V, 3, H, W)), means3D, opac, sc, rot,
                             cov, shs, col, scene.sh_degree, scale_modifier, want_means2D=True, want_conic=True)
        torch.cuda.synchronize()
        grads = {k: v.cpu().numpy() for k, v in g.items()}
    return out, grads, dict(R=R, R_cap=res.R_cap, overflow=int(status[1]) & 0xFFFFFFFF,
                            n_visible=int(status[1]) >> 32)


def rel_err(a, b):
    a = np.asarray(a, np.float64).reshape(-1)
    b = np.asarray(b, np.float64).reshape(-1)
    den = np.abs(b).max()
    if den == 0:
        return float(np.abs(a).max())
    return float(np.abs(a - b).max() / den)


def grad_violation(a, b, rtol=1e-4, floor=1e-6):
    """Per-element gradient bar: max over elements of |a - b| / (rtol |b| + floor max|b|); <= 1 passes.
    Unlike rel_err (max error over max magnitude) a small-magnitude element cannot hide a large relative error
    beyond the absolute floor `floor * max|b|` (fp32 accumulation noise of sums whose terms cancel)."""
    a = np.asarray(a, np.float64).reshape(-1)
    b = np.asarray(b, np.float64).reshape(-1)
    if b.size == 0:
        return 0.0
    den = rtol * np.abs(b) + floor * np.abs(b).max()
    if not np.any(den > 0):
        return float(np.abs(a).max() > 0) * np.inf if np.abs(a).max() > 0 else 0.0
    return float((np.abs(a - b) / np.maximum(den, 1e-300)).max())


class use_library_variant:
    """Route guassianhand_b200 through a build variant (libghr_<name>.so) inside the `with` block."""

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        from guassianhand_b200 import _native, build
        self._native, self._saved = _native, (_native._lib, _native.LIB_PATH)
        path = build.variant_path(self.name)
        if not os.path.exists(path) or build._stale(path):
            build.build(variant=self.name)
        _native._lib, _native.LIB_PATH = None, path
        _native.lib()
        return self

    def __exit__(self, *a):
        self._native._lib, self._native.LIB_PATH = self._saved
