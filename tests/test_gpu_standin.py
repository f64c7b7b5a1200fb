"""The bench-only upstream-STRUCTURED GPU stand-in (baseline_standin/) must compute the same
thing as libghr before its time may stand next to libghr's: identical sorted keys and tile ranges (a global CUB
radix sort against libghr's tile-first binning), the same image, and the same blend-stage gradients
(one CTA per tile + per-thread atomicAdd against the segment-parallel moment backward)."""
import numpy as np
import pytest

import util

pytestmark = pytest.mark.gpu


def _render_both(P, H, W, seed, cuda_device):
    import torch
    from baseline_standin.standin import Standin
    from guassianhand_b200 import api, scenes
    dev = cuda_device
    sc = scenes.two_hand_scene(P, seed=seed)
    cam = scenes.fibonacci_cameras(3, H, W, seed=seed)[1]
    bg = np.array([0.2, 0.5, 0.1], np.float32)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).float().to(dev)
    views = util.gpu_views([cam], bg, dev)
    c = views.cams()
    inp = [t(sc.means3D), t(sc.opacities), t(sc.scales), t(sc.rotations)]
    col = t(sc.colors)
    res = api.forward_raw(c, inp[0], inp[1], inp[2], inp[3], None, None, col, 0, 1.0, want_debug=True)
    lay = res.debug["layout"]
    geom = res.state[lay.off_geom: lay.off_geom + P * 64].view(torch.float32)
    si = Standin(geom, H, W, t(bg))
    img = si.forward()
    dL = t(np.random.default_rng(seed).normal(size=(1, 3, H, W)).astype(np.float32))
    g = api.backward_raw(c, res.state, res.R_cap, dL, inp[0], inp[1], inp[2], inp[3], None, None, col, 0, 1.0,
                         want_means2D=True, want_conic=True)
    sg = si.backward(dL[0].contiguous())
    torch.cuda.synchronize()
    return res, lay, si, img, g, sg


@pytest.mark.parametrize("P,H,W,seed", [(20000, 176, 256, 1), (60000, 334, 512, 2), (5, 40, 56, 3)])
def test_standin_matches_libghr(P, H, W, seed, cuda_device):
    import torch
    res, lay, si, img, g, sg = _render_both(P, H, W, seed, cuda_device)
    R = si.R
    assert R == res.R
    T = ((W + 15) // 16) * ((H + 15) // 16)
    keys = res.debug["keys"][:R]
    assert torch.equal(si.sorted_keys(), keys)                              # global CUB sort == tile-first binning
    ranges = res.state[lay.off_ranges: lay.off_ranges + T * 8].view(torch.int32).view(T, 2)
    nz = ranges[:, 1] > ranges[:, 0]
    assert torch.equal(si.ranges()[nz], ranges[nz])
    assert (img - res.color[0]).abs().max().item() <= 2e-5                  # __expf here, ex2.approx there
    nc = res.state[lay.off_ncontrib: lay.off_ncontrib + H * W * 4].view(torch.int32).view(H, W)
    assert (si.n_contrib() != nc).sum().item() <= max(2, H * W // 20000)
    dm, dc, do, dcol = (x.cpu().numpy() for x in sg)
    for name, a, b in (("mean2D", g["dL_dmeans2D"][0, :, :2].cpu().numpy(), dm),
                       ("conic", g["dL_dconic"][0][:, [0, 1, 3]].cpu().numpy(), dc),
                       ("opacity", g["dL_dopacity"].cpu().numpy().reshape(-1), do),
                       ("colors", g["dL_dcolors"].cpu().numpy(), dcol)):
        assert util.rel_err(a, b) <= 1e-4, name
    si.close()
