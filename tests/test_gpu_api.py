"""The drop-in Python surface on the GPU, driven the way GuassianHand drives it
(/root/reference/tgs/models/renderer_one_shot.py:259-382): zero `means2D` grad holder, keyword
arguments, an RGB render and an all-ones mask render sharing the geometry, both differentiated."""
import numpy as np
import pytest
import torch

from guassianhand_b200 import scenes
import util

pytestmark = pytest.mark.gpu


def _settings(cam, bg, dev, sh_degree=0, debug=False):
    from diff_gaussian_rasterization import GaussianRasterizationSettings
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).float().to(dev)
    # the reference hands over a NON-contiguous view for viewmatrix (w2c.transpose(0,1), :96)
    view = t(cam.viewmatrix.T).transpose(0, 1)
    return GaussianRasterizationSettings(
        image_height=cam.H, image_width=cam.W, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy, bg=t(bg),
        scale_modifier=1.0, viewmatrix=view, projmatrix=t(cam.projmatrix), sh_degree=sh_degree,
        campos=t(cam.campos), prefiltered=False, debug=debug)


def test_reference_call_pattern_rgb_plus_mask(cuda_device):
    from diff_gaussian_rasterization import GaussianRasterizer
    dev = cuda_device
    sc = scenes.two_hand_scene(4000, seed=9)
    cam = scenes.fibonacci_cameras(2, 96, 80, seed=9)[1]
    bg = np.array([0.2, 0.3, 0.4], np.float32)
    leaf = lambda a: torch.from_numpy(a).to(dev).requires_grad_(True)
    xyz, opacity, scaling, rotation, colors = map(leaf, (sc.means3D, sc.opacities, sc.scales, sc.rotations, sc.colors))
    screenspace_points = torch.zeros_like(xyz, requires_grad=True) + 0
    screenspace_points.retain_grad()
    rasterizer = GaussianRasterizer(raster_settings=_settings(cam, bg, dev))
    with torch.autocast(device_type="cuda", dtype=torch.float32):
        rendered_image, radii = rasterizer(means3D=xyz, means2D=screenspace_points, shs=None, colors_precomp=colors,
                                           opacities=opacity, scales=scaling, rotations=rotation, cov3D_precomp=None)
    mask_rasterizer = GaussianRasterizer(raster_settings=_settings(cam, np.zeros(3, np.float32), dev))
    rendered_mask, radii2 = mask_rasterizer(means3D=xyz, means2D=screenspace_points,
                                            colors_precomp=torch.ones_like(xyz), opacities=opacity, scales=scaling,
                                            rotations=rotation, cov3D_precomp=None)
    assert rendered_image.shape == (3, cam.H, cam.W) and radii.dtype == torch.int32 and radii.shape == (sc.P,)
    assert torch.equal(radii, radii2)
    rng = np.random.default_rng(0)
    w_img = torch.from_numpy((rng.normal(size=(3, cam.H, cam.W)) / (cam.H * cam.W)).astype(np.float32)).to(dev)
    w_msk = torch.from_numpy((rng.normal(size=(3, cam.H, cam.W)) / (cam.H * cam.W)).astype(np.float32)).to(dev)
    loss = (rendered_image.permute(1, 2, 0) * w_img.permute(1, 2, 0)).sum() + (rendered_mask * w_msk).sum()
    loss.backward()
    # oracle: the two renders' gradients add
    f1, g1 = util.run_oracle(sc, cam, bg, w_img.cpu().numpy())
    ones = scenes.GaussianScene(**{**sc.__dict__})
    ones.colors = np.ones_like(sc.colors)
    f2, g2 = util.run_oracle(ones, cam, np.zeros(3, np.float32), w_msk.cpu().numpy())
    assert np.array_equal(radii.cpu().numpy(), f1["radii"])
    assert np.abs(rendered_image.detach().cpu().numpy() - f1["out_color"]).max() <= 1e-5
    assert np.abs(rendered_mask.detach().cpu().numpy() - f2["out_color"]).max() <= 1e-5
    tol = 1e-4
    assert util.rel_err(xyz.grad.cpu().numpy(), g1["dL_dmeans3D"].astype(np.float64) + g2["dL_dmeans3D"]) <= tol
    assert util.rel_err(opacity.grad.cpu().numpy().reshape(-1), g1["dL_dopacity"].astype(np.float64) + g2["dL_dopacity"]) <= tol
    assert util.rel_err(scaling.grad.cpu().numpy(), g1["dL_dscales"].astype(np.float64) + g2["dL_dscales"]) <= tol
    assert util.rel_err(rotation.grad.cpu().numpy(), g1["dL_drots"].astype(np.float64) + g2["dL_drots"]) <= tol
    assert util.rel_err(colors.grad.cpu().numpy(), g1["dL_dcolors"]) <= tol
    assert util.rel_err(screenspace_points.grad.cpu().numpy(),
                        g1["dL_dmeans2D"].astype(np.float64) + g2["dL_dmeans2D"]) <= tol
    assert (screenspace_points.grad[:, 2] == 0).all()


def test_sh_path_and_mark_visible(cuda_device):
    from diff_gaussian_rasterization import GaussianRasterizer
    from oracle import oracle_lib as ol
    dev = cuda_device
    sc = scenes.random_scene(1200, seed=6, sh_degree=3)
    cam = scenes.simple_camera(64, 64)
    bg = np.zeros(3, np.float32)
    leaf = lambda a: torch.from_numpy(a).to(dev).requires_grad_(True)
    xyz, opacity, scaling, rotation, shs = map(leaf, (sc.means3D, sc.opacities, sc.scales, sc.rotations, sc.shs))
    m2d = torch.zeros_like(xyz, requires_grad=True)
    r = GaussianRasterizer(raster_settings=_settings(cam, bg, dev, sh_degree=3))
    img, radii = r(means3D=xyz, means2D=m2d, shs=shs, opacities=opacity, scales=scaling, rotations=rotation)
    w = torch.from_numpy(np.random.default_rng(1).normal(size=(3, 64, 64)).astype(np.float32) / 4096).to(dev)
    (img * w).sum().backward()
    f, g = util.run_oracle(sc, cam, bg, w.cpu().numpy())
    assert util.rel_err(shs.grad.cpu().numpy(), g["dL_dsh"]) <= 1e-4
    assert util.rel_err(xyz.grad.cpu().numpy(), g["dL_dmeans3D"]) <= 1e-4
    vis = r.markVisible(xyz.detach())
    assert vis.dtype == torch.bool
    assert np.array_equal(vis.cpu().numpy(), ol.mark_visible(sc.means3D, cam.viewmatrix, cam.projmatrix))


def test_argument_errors_match_upstream(cuda_device):
    from diff_gaussian_rasterization import GaussianRasterizer
    dev = cuda_device
    cam = scenes.simple_camera(32, 32)
    r = GaussianRasterizer(raster_settings=_settings(cam, np.zeros(3, np.float32), dev))
    x = torch.zeros(4, 3, device=dev)
    o = torch.ones(4, 1, device=dev)
    with pytest.raises(Exception, match="excatly one of either SHs or precomputed colors"):
        r(means3D=x, means2D=x, opacities=o, scales=x, rotations=torch.zeros(4, 4, device=dev))
    with pytest.raises(Exception, match="exactly one of either scale/rotation pair or precomputed 3D covariance"):
        r(means3D=x, means2D=x, opacities=o, colors_precomp=x)
    with pytest.raises(RuntimeError, match="CUDA tensors"):
        r(means3D=x.cpu(), means2D=x.cpu(), opacities=o.cpu(), colors_precomp=x.cpu(), scales=x.cpu(),
          rotations=torch.zeros(4, 4))


def test_prefiltered_violation_raises_and_context_survives(cuda_device):
    """upstream's in_frustum prints "Point is filtered although prefiltered is set" and traps when
    settings.prefiltered is set and a point fails the near-plane test; here the forward raises that text (no lost
    CUDA context).  With every point in front of the camera prefiltered=True renders exactly like False."""
    from diff_gaussian_rasterization import GaussianRasterizer
    dev = cuda_device
    sc = scenes.two_hand_scene(2000, seed=3)
    cam = scenes.fibonacci_cameras(2, 64, 64, seed=3)[0]
    rs = _settings(cam, np.zeros(3, np.float32), dev)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).float().to(dev)
    kw = dict(means2D=torch.zeros(sc.P, 3, device=dev), opacities=t(sc.opacities), colors_precomp=t(sc.colors),
              scales=t(sc.scales), rotations=t(sc.rotations))
    xyz = t(sc.means3D)
    img0, rad0 = GaussianRasterizer(raster_settings=rs)(means3D=xyz, **kw)
    img1, rad1 = GaussianRasterizer(raster_settings=rs._replace(prefiltered=True))(means3D=xyz, **kw)
    assert torch.equal(img0, img1) and torch.equal(rad0, rad1)
    behind = xyz.clone()
    behind[7] = t(cam.campos) - 5.0 * (xyz.mean(0) - t(cam.campos))         # one point behind the camera
    with pytest.raises(RuntimeError, match="Point is filtered although prefiltered is set"):
        GaussianRasterizer(raster_settings=rs._replace(prefiltered=True))(means3D=behind, **kw)
    img2, _ = GaussianRasterizer(raster_settings=rs)(means3D=behind, **kw)   # the context is alive, the point culled
    assert torch.isfinite(img2).all()


def test_batched_views_autograd_equals_view_loop(cuda_device):
    from guassianhand_b200 import rasterize_views
    dev = cuda_device
    sc = scenes.two_hand_scene(3000, seed=4)
    cams = scenes.fibonacci_cameras(4, 80, 96, seed=4)
    bg = np.array([0.0, 0.1, 0.2], np.float32)
    views = util.gpu_views(cams, bg, dev)
    leaf = lambda a: torch.from_numpy(a).to(dev).requires_grad_(True)
    xyz, opacity, scaling, rotation, colors = map(leaf, (sc.means3D, sc.opacities, sc.scales, sc.rotations, sc.colors))
    imgs, radii = rasterize_views(xyz, opacity, views, colors_precomp=colors, scales=scaling, rotations=rotation)
    w = torch.from_numpy((np.random.default_rng(2).normal(size=(4, 3, 80, 96)) / 7680).astype(np.float32)).to(dev)
    (imgs * w).sum().backward()
    tot = {}
    for v, cam in enumerate(cams):
        f, g = util.run_oracle(sc, cam, bg, w[v].cpu().numpy())
        assert np.abs(imgs[v].detach().cpu().numpy() - f["out_color"]).max() <= 1e-5
        assert np.array_equal(radii[v].cpu().numpy(), f["radii"])
        for k, a in g.items():
            tot[k] = tot.get(k, 0) + a.astype(np.float64)
    assert util.rel_err(xyz.grad.cpu().numpy(), tot["dL_dmeans3D"]) <= 1e-4
    assert util.rel_err(colors.grad.cpu().numpy(), tot["dL_dcolors"]) <= 1e-4
    assert util.rel_err(rotation.grad.cpu().numpy(), tot["dL_drots"]) <= 1e-4


def test_fused_mask_equals_second_render(cuda_device):
    """the coverage mask of the SAME pass (1 - T_final) must equal the
    reference's second render (colors = 1, bg = 0, renderer_one_shot.py:353-380) to 1e-5, and one
    backward must deliver the sum of both renders' gradients to 1e-4."""
    from diff_gaussian_rasterization import GaussianRasterizer
    from guassianhand_b200 import rasterize_views
    dev = cuda_device
    sc = scenes.two_hand_scene(4000, seed=19)
    cams = scenes.fibonacci_cameras(3, 96, 80, seed=19)
    bg = np.array([0.2, 0.3, 0.4], np.float32)
    rng = np.random.default_rng(5)
    w_img = (rng.normal(size=(3, 3, 96, 80)) / 7680).astype(np.float32)
    w_msk = (rng.normal(size=(3, 96, 80)) / 7680).astype(np.float32)
    leaf = lambda a: torch.from_numpy(a).to(dev).requires_grad_(True)
    # single-view drop-in object
    xyz, opacity, scaling, rotation, colors = map(leaf, (sc.means3D, sc.opacities, sc.scales, sc.rotations, sc.colors))
    m2d = torch.zeros_like(xyz, requires_grad=True)
    r = GaussianRasterizer(raster_settings=_settings(cams[0], bg, dev))
    img, radii, mask = r.forward_with_mask(means3D=xyz, means2D=m2d, colors_precomp=colors, opacities=opacity,
                                           scales=scaling, rotations=rotation)
    ((img * torch.from_numpy(w_img[0]).to(dev)).sum() + (mask * torch.from_numpy(w_msk[0]).to(dev)).sum()).backward()
    f1, g1 = util.run_oracle(sc, cams[0], bg, w_img[0])
    ones = scenes.GaussianScene(**{**sc.__dict__})
    ones.colors = np.ones_like(sc.colors)
    # the mask render's three channels are identical; dL/dmask spreads to one channel of the oracle call
    dLm = np.stack([w_msk[0], np.zeros_like(w_msk[0]), np.zeros_like(w_msk[0])])
    f2, g2 = util.run_oracle(ones, cams[0], np.zeros(3, np.float32), dLm)
    assert np.abs(mask.detach().cpu().numpy() - f2["out_color"][0]).max() <= 1e-5
    assert np.abs(img.detach().cpu().numpy() - f1["out_color"]).max() <= 1e-5
    for t, k in ((xyz, "dL_dmeans3D"), (scaling, "dL_dscales"), (rotation, "dL_drots")):
        assert util.rel_err(t.grad.cpu().numpy(), g1[k].astype(np.float64) + g2[k]) <= 1e-4, k
    assert util.rel_err(opacity.grad.cpu().numpy().reshape(-1), g1["dL_dopacity"].astype(np.float64) + g2["dL_dopacity"]) <= 1e-4
    assert util.rel_err(m2d.grad.cpu().numpy(), g1["dL_dmeans2D"].astype(np.float64) + g2["dL_dmeans2D"]) <= 1e-4
    assert util.rel_err(colors.grad.cpu().numpy(), g1["dL_dcolors"]) <= 1e-4
    # batched entry
    xyz2, opacity2, scaling2, rotation2, colors2 = map(leaf, (sc.means3D, sc.opacities, sc.scales, sc.rotations, sc.colors))
    views = util.gpu_views(cams, bg, dev)
    imgs, masks, _ = rasterize_views(xyz2, opacity2, views, colors_precomp=colors2, scales=scaling2,
                                     rotations=rotation2, return_mask=True)
    ((imgs * torch.from_numpy(w_img).to(dev)).sum() + (masks * torch.from_numpy(w_msk).to(dev)).sum()).backward()
    tot = 0
    for v, cam in enumerate(cams):
        _, ga = util.run_oracle(sc, cam, bg, w_img[v])
        fb, gb = util.run_oracle(ones, cam, np.zeros(3, np.float32),
                                 np.stack([w_msk[v], np.zeros_like(w_msk[v]), np.zeros_like(w_msk[v])]))
        assert np.abs(masks[v].detach().cpu().numpy() - fb["out_color"][0]).max() <= 1e-5
        tot = tot + ga["dL_dmeans3D"].astype(np.float64) + gb["dL_dmeans3D"]
    assert util.rel_err(xyz2.grad.cpu().numpy(), tot) <= 1e-4


def test_overlapped_view_groups_equal_single_chain(cuda_device):
    """rasterize_views(overlap=G) and dist.fit_step_grads(overlap=G) only change the schedule (view groups
    on concurrent streams): images, masks and radii are bit-identical, gradients equal up to the order
    of the sum over views."""
    from guassianhand_b200 import rasterize_views
    from guassianhand_b200.dist import GraphedFitStep, PackedGrads, fit_step_grads
    dev = cuda_device
    sc = scenes.two_hand_scene(5000, seed=11)
    cams = scenes.fibonacci_cameras(5, 96, 112, seed=11)
    bg = np.stack([np.array([0.1 * v, 0.2, 0.3], np.float32) for v in range(5)])     # per-view background
    views = util.gpu_views(cams, bg, dev)
    t = lambda a: torch.from_numpy(a).to(dev)
    w = t((np.random.default_rng(3).normal(size=(5, 3, 96, 112)) / 10752).astype(np.float32))
    wm = t((np.random.default_rng(4).normal(size=(5, 96, 112)) / 10752).astype(np.float32))
    out = {}
    for G in (1, 3, 5):
        leaf = lambda a: t(a).requires_grad_(True)
        xyz, opacity, scaling, rotation, colors = map(leaf, (sc.means3D, sc.opacities, sc.scales, sc.rotations, sc.colors))
        imgs, mask, radii = rasterize_views(xyz, opacity, views, colors_precomp=colors, scales=scaling,
                                            rotations=rotation, return_mask=True, overlap=G)
        ((imgs * w).sum() + (mask * wm).sum()).backward()
        torch.cuda.synchronize()
        out[G] = (imgs.detach().clone(), mask.detach().clone(), radii.clone(),
                  [x.grad.clone() for x in (xyz, opacity, scaling, rotation, colors)])
    for G in (3, 5):
        assert torch.equal(out[G][0], out[1][0]) and torch.equal(out[G][1], out[1][1]) and torch.equal(out[G][2], out[1][2])
        for a, b in zip(out[G][3], out[1][3]):
            assert util.rel_err(a.cpu().numpy(), b.cpu().numpy().astype(np.float64)) <= 1e-5

    gauss = dict(means3D=t(sc.means3D), opacities=t(sc.opacities), scales=t(sc.scales), rotations=t(sc.rotations),
                 colors_precomp=t(sc.colors))
    g1, g2, g3 = (PackedGrads(sc.P, 0, device=dev) for _ in range(3))
    r1 = fit_step_grads(gauss, views, w, g1)
    r2 = fit_step_grads(gauss, views, w, g2, overlap=2)
    assert r2.R == r1.R and torch.equal(r2.color, r1.color)
    assert util.rel_err(g2.flat.cpu().numpy(), g1.flat.cpu().numpy().astype(np.float64)) <= 1e-5
    caps = [int(x.R * 1.25) + 1024 for x in r2.results]
    step = GraphedFitStep(gauss, views, w, g3, R_cap=caps, overlap=2)
    step.replay()
    R, overflow = step.status()
    assert R == r1.R and not overflow
    assert util.rel_err(g3.flat.cpu().numpy(), g1.flat.cpu().numpy().astype(np.float64)) <= 1e-5


def test_odd_gaussian_count_packed_grads_and_graph(cuda_device):
    """One reference hand has 49,281 points (odd): every PackedGrads segment must stay 16-byte aligned (the
    backward writes rotations with 128-bit stores), eagerly and through the CUDA graph; the graph owns its
    scratch, so a later call that regrows the shared workspace must not disturb a replay."""
    from guassianhand_b200 import api
    from guassianhand_b200.dist import GraphedFitStep, PackedGrads, fit_step_grads
    dev = cuda_device
    P = 3001
    sc = scenes.two_hand_scene(P, seed=21)
    cams = scenes.fibonacci_cameras(3, 80, 96, seed=21)
    bg = np.array([0.1, 0.0, 0.2], np.float32)
    views = util.gpu_views(cams, bg, dev)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).float().to(dev)
    w = t((np.random.default_rng(6).normal(size=(3, 3, 80, 96)) / 7680).astype(np.float32))
    # rotations as a 4-byte-aligned view into a packed parameter buffer (flat[6P:10P] with odd P)
    flat = torch.zeros(14 * P + 1, device=dev)
    rot_view = flat[6 * P + 1: 10 * P + 1].view(P, 4)
    rot_view.copy_(t(sc.rotations))
    assert rot_view.data_ptr() % 16 != 0
    gauss = dict(means3D=t(sc.means3D), opacities=t(sc.opacities), scales=t(sc.scales), rotations=rot_view,
                 colors_precomp=t(sc.colors))
    g1, g2 = PackedGrads(P, 0, device=dev), PackedGrads(P, 0, device=dev)
    r1 = fit_step_grads(gauss, views, w, g1)
    tot = {}
    for v, cam in enumerate(cams):
        _, g = util.run_oracle(sc, cam, bg, w[v].cpu().numpy())
        for k, a in g.items():
            tot[k] = tot.get(k, 0) + a.astype(np.float64)
    for name, key in (("dL_dmeans3D", "dL_dmeans3D"), ("dL_dscales", "dL_dscales"), ("dL_drotations", "dL_drots"),
                      ("dL_dcolors", "dL_dcolors")):
        assert util.rel_err(g1.views()[name].cpu().numpy(), tot[key]) <= 1e-4, name
    caps = [int(x.R * 1.25) + 1024 for x in fit_step_grads(gauss, views, w, g2, overlap=2).results]
    step = GraphedFitStep(gauss, views, w, g2, R_cap=caps, overlap=2)
    step.replay()
    torch.cuda.synchronize()
    ref = g2.flat.clone()
    # force the shared per-stream workspaces to regrow, then replay again: same result
    big = scenes.two_hand_scene(20000, seed=22)
    bviews = util.gpu_views(scenes.fibonacci_cameras(2, 160, 160, seed=22), bg, dev)
    for st in [torch.cuda.current_stream()] + api._streams(torch.device(dev), 1):
        with torch.cuda.stream(st):
            api.forward_raw(bviews.cams(), t(big.means3D), t(big.opacities), t(big.scales), t(big.rotations), None, None,
                            t(big.colors), 0, 1.0)
    torch.cuda.synchronize()
    g2.zero_()
    step.replay()
    torch.cuda.synchronize()
    assert not step.status()[1]
    assert util.rel_err(g2.flat.cpu().numpy(), ref.cpu().numpy().astype(np.float64)) <= 1e-5
    assert util.rel_err(g2.flat.cpu().numpy(), g1.flat.cpu().numpy().astype(np.float64)) <= 1e-5


def test_misaligned_rotations_through_drop_in(cuda_device):
    """A contiguous [P,4] rotation view that is only 4-byte aligned must render (cloned to an aligned buffer)
    and receive its gradient."""
    from diff_gaussian_rasterization import GaussianRasterizer
    dev = cuda_device
    P = 777
    sc = scenes.two_hand_scene(P, seed=23)
    cam = scenes.fibonacci_cameras(1, 64, 64, seed=23)[0]
    bg = np.zeros(3, np.float32)
    leaf = lambda a: torch.from_numpy(a).to(dev).requires_grad_(True)
    xyz, opacity, scaling, colors = map(leaf, (sc.means3D, sc.opacities, sc.scales, sc.colors))
    packed = torch.zeros(4 * P + 1, device=dev)
    packed[1:] = torch.from_numpy(sc.rotations).to(dev).reshape(-1)
    packed.requires_grad_(True)
    rotation = packed[1:].view(P, 4)
    assert rotation.data_ptr() % 16 != 0
    m2d = torch.zeros_like(xyz, requires_grad=True)
    r = GaussianRasterizer(raster_settings=_settings(cam, bg, dev))
    img, _ = r(means3D=xyz, means2D=m2d, colors_precomp=colors, opacities=opacity, scales=scaling, rotations=rotation)
    wgt = torch.from_numpy((np.random.default_rng(8).normal(size=(3, 64, 64)) / 4096).astype(np.float32)).to(dev)
    (img * wgt).sum().backward()
    f, g = util.run_oracle(sc, cam, bg, wgt.cpu().numpy())
    assert np.abs(img.detach().cpu().numpy() - f["out_color"]).max() <= 1e-5
    assert util.rel_err(packed.grad[1:].view(P, 4).cpu().numpy(), g["dL_drots"]) <= 1e-4


def test_device_side_cameras_from_w2c(cuda_device):
    """ViewBatch.from_w2c (ghr_cameras_from_w2c, one launch, no host sync) against the host restatement of the
    reference's Camera.from_w2c / getProjectionMatrix_refine / intrinsic_to_fov + math.tan
    (scenes.camera_from_w2c <- renderer_one_shot.py:61-112, :278-279): matrices to 1 ulp, and the same scene
    rendered through both camera sets gives identical radii and images."""
    from guassianhand_b200 import api
    dev = cuda_device
    H, W, V = 96, 112, 6
    rng = np.random.default_rng(12)
    fx = 1300.0 * W / 334.0
    w2cs, Ks, host = [], [], []
    for v in range(V):
        eye = rng.normal(size=3)
        eye = eye / np.linalg.norm(eye) * rng.uniform(0.8, 1.3)
        eye[2] = -abs(eye[2]) - 0.3
        w2c = scenes.look_at_w2c(eye, rng.normal(size=3) * 0.02)
        K = np.array([[fx * rng.uniform(0.9, 1.1), rng.uniform(-0.5, 0.5), W / 2 + rng.uniform(-5, 5)],
                      [0, fx * rng.uniform(0.9, 1.1), H / 2 + rng.uniform(-5, 5)], [0, 0, 1]])
        w2cs.append(w2c.astype(np.float32))
        Ks.append(K.astype(np.float32))
        host.append(scenes.camera_from_w2c(w2c, K, H, W))
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).float().to(dev)
    bg = np.array([0.1, 0.2, 0.3], np.float32)
    vb = api.ViewBatch.from_w2c(t(np.stack(w2cs)), t(np.stack(Ks)), H, W, t(bg))
    torch.cuda.synchronize()
    ulp = lambda a, b: np.abs(a.astype(np.float64) - b.astype(np.float64)) / np.maximum(np.spacing(np.abs(b).astype(np.float32)), 1e-45)
    for v in range(V):
        assert np.array_equal(vb.viewmatrix[v].cpu().numpy(), host[v].viewmatrix)
        pm, hm = vb.projmatrix[v].cpu().numpy(), host[v].projmatrix
        assert np.abs(pm - hm).max() <= 4e-7 * np.abs(hm).max()           # fp32 4-term dot products
        assert ulp(vb.campos[v].cpu().numpy(), host[v].campos).max() <= 1.0
        tf = vb.tanfov[v].cpu().numpy()
        assert ulp(tf, np.array([host[v].tanfovx, host[v].tanfovy], np.float32)).max() <= 2.0
    # end to end: identical integers, images to 1e-5
    sc = scenes.two_hand_scene(3000, seed=5)
    tt = lambda a: torch.from_numpy(np.ascontiguousarray(a)).float().to(dev)
    args = (tt(sc.means3D), tt(sc.opacities), tt(sc.scales), tt(sc.rotations), None, None, tt(sc.colors), 0, 1.0)
    r_dev = api.forward_raw(vb.cams(), *args)
    r_host = api.forward_raw(util.gpu_views(host, bg, dev).cams(), *args)
    torch.cuda.synchronize()
    assert (r_dev.radii != r_host.radii).float().mean().item() <= 1e-3     # a 1-ulp tanfov may move a ceil()
    assert (r_dev.color - r_host.color).abs().max().item() <= 2e-3
    same = (r_dev.radii == r_host.radii).all().item()
    if same:
        assert (r_dev.color - r_host.color).abs().max().item() <= 1e-5


def test_geometry_cache_second_call_reuses_binning(cuda_device):
    """The reference renders every view twice with identical geometry (RGB, then colours = 1 / bg = 0 / sh_degree 0,
    renderer_one_shot.py:338-346, :372-379), passing the SAME tensor objects for geometry and camera.  The second
    call must take the geometry-reuse path, give bit-identical outputs and gradients to a call that does not,
    and an in-place change of a geometry tensor must invalidate the cache."""
    from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer
    from guassianhand_b200 import api
    dev = cuda_device
    sc = scenes.two_hand_scene(5000, seed=33)
    cam = scenes.fibonacci_cameras(2, 128, 112, seed=33)[1]
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).float().to(dev)
    view, proj, campos = t(cam.viewmatrix.T).transpose(0, 1), t(cam.projmatrix), t(cam.campos)   # shared by both calls

    def settings(bg):
        return GaussianRasterizationSettings(
            image_height=cam.H, image_width=cam.W, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy, bg=bg, scale_modifier=1.0,
            viewmatrix=view, projmatrix=proj, sh_degree=0, campos=campos, prefiltered=False, debug=False)
    w_img = t((np.random.default_rng(0).normal(size=(3, cam.H, cam.W)) / (cam.H * cam.W)).astype(np.float32))
    w_msk = t((np.random.default_rng(1).normal(size=(3, cam.H, cam.W)) / (cam.H * cam.W)).astype(np.float32))

    def pair(clear_between):
        leaf = lambda a: torch.from_numpy(a).to(dev).requires_grad_(True)
        xyz, opacity, scaling, rotation, colors = map(leaf, (sc.means3D, sc.opacities, sc.scales, sc.rotations, sc.colors))
        m2d = torch.zeros_like(xyz, requires_grad=True)
        api._GeomCache.entries.clear()
        img, radii = GaussianRasterizer(raster_settings=settings(t(np.array([0.2, 0.3, 0.4], np.float32))))(
            means3D=xyz, means2D=m2d, shs=None, colors_precomp=colors, opacities=opacity, scales=scaling,
            rotations=rotation, cov3D_precomp=None)
        if clear_between:
            api._GeomCache.entries.clear()
        msk, radii2 = GaussianRasterizer(raster_settings=settings(torch.zeros(3, device=dev)))(
            means3D=xyz, means2D=m2d, colors_precomp=torch.ones_like(xyz), opacities=opacity, scales=scaling,
            rotations=rotation, cov3D_precomp=None)
        ((img * w_img).sum() + (msk * w_msk).sum()).backward()
        torch.cuda.synchronize()
        return img.detach(), msk.detach(), radii, radii2, [x.grad.clone() for x in (xyz, opacity, scaling, rotation, colors, m2d)]

    h0 = api._GeomCache.hits
    a = pair(clear_between=False)
    assert api._GeomCache.hits == h0 + 1                                 # the mask call reused the RGB call's binning
    b = pair(clear_between=True)
    assert api._GeomCache.hits == h0 + 1
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and torch.equal(a[2], b[2]) and torch.equal(a[3], b[3])
    for ga, gb in zip(a[4], b[4]):
        assert util.rel_err(ga.cpu().numpy(), gb.cpu().numpy().astype(np.float64)) <= 1e-5    # atomics order only
    # oracle: the mask render is what a stand-alone render of colours = 1 gives
    ones = scenes.GaussianScene(**{**sc.__dict__})
    ones.colors = np.ones_like(sc.colors)
    f2, _ = util.run_oracle(ones, cam, np.zeros(3, np.float32))
    assert np.abs(a[1].cpu().numpy() - f2["out_color"]).max() <= 1e-5
    # in-place modification of a geometry tensor between the calls: version bump -> no reuse
    xyz = t(sc.means3D)
    opacity, scaling, rotation, colors = t(sc.opacities), t(sc.scales), t(sc.rotations), t(sc.colors)
    m2d = torch.zeros_like(xyz)
    api._GeomCache.entries.clear()
    r = GaussianRasterizer(raster_settings=settings(torch.zeros(3, device=dev)))
    r(means3D=xyz, means2D=m2d, colors_precomp=colors, opacities=opacity, scales=scaling, rotations=rotation)
    xyz.add_(0.01)
    h1 = api._GeomCache.hits
    img2, _ = r(means3D=xyz, means2D=m2d, colors_precomp=colors, opacities=opacity, scales=scaling, rotations=rotation)
    assert api._GeomCache.hits == h1
    moved = scenes.GaussianScene(**{**sc.__dict__})
    moved.means3D = (sc.means3D + np.float32(0.01)).astype(np.float32)
    f3, _ = util.run_oracle(moved, cam, np.zeros(3, np.float32))
    assert np.abs(img2.cpu().numpy() - f3["out_color"]).max() <= 1e-5
