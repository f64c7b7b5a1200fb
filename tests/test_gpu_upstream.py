"""Optional cross-check against the REAL upstream extension (pip diff-gaussian-rasterization, the package
/root/reference/tgs/models/renderer_one_shot.py:3 imports).  It is not vendored in the reference and not
installable here (no network), so this test is skipped unless a built copy is found under baseline/_ref
(the driver's install location) -- everywhere else parity is pinned to the oracle only ("parity unpinned",
DESIGN.md §3).  When it runs it is the strongest check in the suite: same inputs through both extensions."""
import glob
import importlib.util
import os
import sys

import numpy as np
import pytest
import torch

from guassianhand_b200 import scenes

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load_upstream():
    base = os.path.join(ROOT, "baseline", "_ref")
    hits = glob.glob(os.path.join(base, "**", "diff_gaussian_rasterization", "__init__.py"), recursive=True)
    hits = [h for h in hits if glob.glob(os.path.join(os.path.dirname(h), "_C*.so"))]
    if not hits:
        return None
    spec = importlib.util.spec_from_file_location("upstream_dgr", hits[0], submodule_search_locations=[os.path.dirname(hits[0])])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["upstream_dgr"] = mod
    spec.loader.exec_module(mod)
    return mod


def test_against_upstream_extension(cuda_device):
    up = _load_upstream()
    if up is None:
        pytest.skip("upstream diff_gaussian_rasterization is not installed under baseline/_ref")
    import diff_gaussian_rasterization as ours
    dev = cuda_device
    sc = scenes.two_hand_scene(98562, seed=0)
    cam = scenes.fibonacci_cameras(2, 256, 256, seed=0)[0]
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).float().to(dev)
    w = t((np.random.default_rng(3).normal(size=(3, 256, 256)) / 65536).astype(np.float32))
    results = {}
    for name, mod in (("upstream", up), ("ours", ours)):
        leafs = [t(x).requires_grad_(True) for x in (sc.means3D, sc.opacities, sc.scales, sc.rotations, sc.colors)]
        xyz, op, scl, rot, col = leafs
        m2d = torch.zeros_like(xyz, requires_grad=True)
        rs = mod.GaussianRasterizationSettings(
            image_height=cam.H, image_width=cam.W, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy, bg=torch.zeros(3, device=dev),
            scale_modifier=1.0, viewmatrix=t(cam.viewmatrix), projmatrix=t(cam.projmatrix), sh_degree=0,
            campos=t(cam.campos), prefiltered=False, debug=False)
        img, radii = mod.GaussianRasterizer(raster_settings=rs)(
            means3D=xyz, means2D=m2d, shs=None, colors_precomp=col, opacities=op, scales=scl, rotations=rot,
            cov3D_precomp=None)
        (img * w).sum().backward()
        results[name] = (img.detach().cpu().numpy(), radii.cpu().numpy(), [x.grad.cpu().numpy() for x in leafs + [m2d]])
    (iu, ru, gu), (io, ro, go) = results["upstream"], results["ours"]
    assert np.array_equal(ru, ro)                                   # radii bit-exact
    # expected: <= 1e-5 except on the few pixels where ex2.approx flips an alpha >= 1/255 decision
    assert (np.abs(iu - io).max(axis=0) > 1e-5).sum() <= 5
    for a, b in zip(gu, go):
        assert np.abs(a - b).max() <= 1e-4 * np.abs(a).max()
