"""Pure-PyTorch (CPU, fp32, autograd) restatement of the rasterizer behind
`diff_gaussian_rasterization` -- the *second, independent* oracle.  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED (see oracle/gs_oracle.c header): the reference holds no source, tests or golden
vectors for this path; the algorithm restated here is SURVEY.md Appendix A.  This file is
written against the math (autograd supplies every gradient), gs_oracle.c against the analytic
backward formulas; tests/test_oracle_cross.py requires the two to agree, which is what pins
the oracle.  It is also BASELINE.json config 1 ("pure-PyTorch CPU re-expression").

Inputs mirror GaussianRasterizer.forward (renderer_one_shot.py:338-346) and the settings tuple
(:281-294).  `means2D` enters as a zero NDC-space offset so autograd reproduces the NDC-unit
dL/dmeans2D of the extension (A.6: pixel gradient x 0.5*W, 0.5*H).
"""
import math
from typing import Optional

import torch

C0 = 0.28209479177387814
C1 = 0.4886025119029199
C2 = [1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792,
      0.5462742152960396]
C3 = [-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154,
      -0.4570457994644658, 1.445305721320277, -0.5900435899266435]


def _quat_to_rot(q):
    r, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = torch.stack([
        1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
        2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
        2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], dim=1)
    return R.view(-1, 3, 3)


def _sh_color(deg, shs, dirs):
    x, y, z = dirs[:, 0:1], dirs[:, 1:2], dirs[:, 2:3]
    res = C0 * shs[:, 0]
    if deg > 0:
        res = res - C1 * y * shs[:, 1] + C1 * z * shs[:, 2] - C1 * x * shs[:, 3]
        if deg > 1:
            xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
            res = (res + C2[0] * xy * shs[:, 4] + C2[1] * yz * shs[:, 5]
                   + C2[2] * (2 * zz - xx - yy) * shs[:, 6] + C2[3] * xz * shs[:, 7]
                   + C2[4] * (xx - yy) * shs[:, 8])
            if deg > 2:
                res = (res + C3[0] * y * (3 * xx - yy) * shs[:, 9] + C3[1] * xy * z * shs[:, 10]
                       + C3[2] * y * (4 * zz - xx - yy) * shs[:, 11]
                       + C3[3] * z * (2 * zz - 3 * xx - 3 * yy) * shs[:, 12]
                       + C3[4] * x * (4 * zz - xx - yy) * shs[:, 13]
                       + C3[5] * z * (xx - yy) * shs[:, 14] + C3[6] * x * (xx - 3 * yy) * shs[:, 15])
    return res + 0.5


def preprocess(means3D, means2D, opacities, scales, rotations, cov3D_precomp, shs, colors_precomp, *,
               H, W, tanfovx, tanfovy, viewmatrix, projmatrix, campos, sh_degree, scale_modifier):
    """A.2 + A.3.  Returns dict of per-Gaussian tensors (differentiable where upstream is)."""
    P = means3D.shape[0]
    V = viewmatrix.reshape(4, 4)     # V[c, r]: flat index 4c + r
    PV = projmatrix.reshape(4, 4)
    ones = torch.ones(P, 1, dtype=means3D.dtype)
    ph = torch.cat([means3D, ones], dim=1)
    p_view = ph @ V                   # row-vector convention: [x y z 1] @ (w2c^T)
    p_hom = ph @ PV
    depth = p_view[:, 2]
    p_w = 1.0 / (p_hom[:, 3] + 1e-7)
    ndc = p_hom[:, :2] * p_w[:, None]
    if means2D is not None:
        ndc = ndc + means2D[:, :2]
    if cov3D_precomp is not None:
        c = cov3D_precomp
        Sigma = torch.stack([c[:, 0], c[:, 1], c[:, 2], c[:, 1], c[:, 3], c[:, 4], c[:, 2], c[:, 4], c[:, 5]],
                            dim=1).view(-1, 3, 3)
    else:
        R = _quat_to_rot(rotations)
        A = R * (scale_modifier * scales)[:, None, :]
        Sigma = A @ A.transpose(1, 2)
    fx = W / (2.0 * tanfovx)
    fy = H / (2.0 * tanfovy)
    limx, limy = 1.3 * tanfovx, 1.3 * tanfovy
    tz = p_view[:, 2]
    txtz = p_view[:, 0] / tz
    tytz = p_view[:, 1] / tz
    # A.7: where the clamp is active the extension passes no gradient to t.x / t.y
    tx = torch.where((txtz < -limx) | (txtz > limx), (txtz.clamp(-limx, limx) * tz).detach(), p_view[:, 0])
    ty = torch.where((tytz < -limy) | (tytz > limy), (tytz.clamp(-limy, limy) * tz).detach(), p_view[:, 1])
    zeros = torch.zeros_like(tz)
    J = torch.stack([fx / tz, zeros, -(fx * tx) / (tz * tz),
                     zeros, fy / tz, -(fy * ty) / (tz * tz)], dim=1).view(-1, 2, 3)
    Wm = V[:3, :3].t()                # Wm[r, c] = flat[4c + r]
    T = J @ Wm                        # [P,2,3]
    cov = T @ Sigma @ T.transpose(1, 2)
    a = cov[:, 0, 0] + 0.3
    b = cov[:, 0, 1]
    c_ = cov[:, 1, 1] + 0.3
    det = a * c_ - b * b
    det_safe = torch.where(det == 0, torch.ones_like(det), det)
    conic = torch.stack([c_ / det_safe, -b / det_safe, a / det_safe], dim=1)
    mid = 0.5 * (a + c_)
    sq = torch.sqrt(torch.clamp_min(mid * mid - det, 0.1))
    lam = torch.maximum(mid + sq, mid - sq)
    radius = torch.ceil(3.0 * torch.sqrt(lam)).to(torch.int32)
    pix = torch.stack([((ndc[:, 0] + 1.0) * W - 1.0) * 0.5, ((ndc[:, 1] + 1.0) * H - 1.0) * 0.5], dim=1)
    gx, gy = (W + 15) // 16, (H + 15) // 16
    pd = pix.detach()
    rf = radius.to(pd.dtype)
    minx = ((pd[:, 0] - rf) / 16.0).to(torch.int32).clamp(0, gx)
    miny = ((pd[:, 1] - rf) / 16.0).to(torch.int32).clamp(0, gy)
    maxx = ((pd[:, 0] + rf + 16.0 - 1.0) / 16.0).to(torch.int32).clamp(0, gx)
    maxy = ((pd[:, 1] + rf + 16.0 - 1.0) / 16.0).to(torch.int32).clamp(0, gy)
    tiles = (maxx - minx) * (maxy - miny)
    visible = (depth.detach() > 0.2) & (det.detach() != 0) & (tiles > 0)
    if colors_precomp is not None:
        rgb = colors_precomp
        clamped = torch.zeros(P, 3, dtype=torch.bool)
    else:
        dirs = means3D - campos[None, :]
        dirs = dirs / dirs.norm(dim=1, keepdim=True)
        raw = _sh_color(sh_degree, shs, dirs)
        clamped = raw.detach() < 0
        rgb = torch.clamp_min(raw, 0.0)
    radius = torch.where(visible, radius, torch.zeros_like(radius))
    tiles = torch.where(visible, tiles, torch.zeros_like(tiles))
    return dict(depth=depth.detach(), radius=radius, pix=pix, conic=conic, rgb=rgb, clamped=clamped,
                tiles=tiles, rect=(minx, miny, maxx, maxy), visible=visible, cov3D=Sigma)


def bin_tiles(pre, H, W):
    """A.4: returns (keys_sorted int64, point_list int64, ranges int64[T,2])."""
    gx, gy = (W + 15) // 16, (H + 15) // 16
    minx, miny, maxx, maxy = pre["rect"]
    vis = pre["visible"].nonzero().flatten()
    keys, vals = [], []
    dbits = pre["depth"].view(torch.int32).to(torch.int64) & 0xFFFFFFFF
    for i in vis.tolist():
        ys = torch.arange(int(miny[i]), int(maxy[i]))
        xs = torch.arange(int(minx[i]), int(maxx[i]))
        t = (ys[:, None] * gx + xs[None, :]).flatten().to(torch.int64)
        keys.append((t << 32) | dbits[i])
        vals.append(torch.full_like(t, i))
    if keys:
        keys = torch.cat(keys)
        vals = torch.cat(vals)
        order = torch.sort(keys, stable=True).indices
        keys, vals = keys[order], vals[order]
    else:
        keys = torch.zeros(0, dtype=torch.int64)
        vals = torch.zeros(0, dtype=torch.int64)
    ranges = torch.zeros(gx * gy, 2, dtype=torch.int64)
    if keys.numel():
        tile = keys >> 32
        uniq, counts = torch.unique_consecutive(tile, return_counts=True)
        ends = torch.cumsum(counts, 0)
        ranges[uniq, 0] = ends - counts
        ranges[uniq, 1] = ends
    return keys, vals, ranges


def render(pre, point_list, ranges, bg, H, W):
    """A.5, one tile at a time, vectorised over (256 pixels) x (tile list)."""
    gx, gy = (W + 15) // 16, (H + 15) // 16
    out = torch.zeros(3, H, W, dtype=pre["pix"].dtype)
    out = out + bg.view(3, 1, 1) * 1.0        # T=1 where nothing is blended
    n_contrib = torch.zeros(H, W, dtype=torch.int64)
    final_T = torch.ones(H, W, dtype=pre["pix"].dtype)
    pieces = []
    for t in range(gx * gy):
        s, e = int(ranges[t, 0]), int(ranges[t, 1])
        if e <= s:
            continue
        ty, tx = divmod(t, gx)
        y0, x0 = ty * 16, tx * 16
        y1, x1 = min(y0 + 16, H), min(x0 + 16, W)
        ys, xs = torch.meshgrid(torch.arange(y0, y1), torch.arange(x0, x1), indexing="ij")
        pxf = xs.flatten().to(pre["pix"].dtype)
        pyf = ys.flatten().to(pre["pix"].dtype)
        ids = point_list[s:e]
        xy = pre["pix"][ids]
        con = pre["conic"][ids]
        # opacity is carried in pre["opacity"]
        op = pre["opacity"][ids]
        col = pre["rgb"][ids]
        dx = xy[None, :, 0] - pxf[:, None]
        dy = xy[None, :, 1] - pyf[:, None]
        power = -0.5 * (con[None, :, 0] * dx * dx + con[None, :, 2] * dy * dy) - con[None, :, 1] * dx * dy
        G = torch.exp(power)
        oG = op[None, :] * G
        alpha = oG + (torch.clamp_max(oG, 0.99) - oG).detach()     # straight-through clamp (A.6)
        valid = (power.detach() <= 0) & (alpha.detach() >= 1.0 / 255.0)
        a = torch.where(valid, alpha, torch.zeros_like(alpha))
        one_m = 1.0 - a
        Tincl = torch.cumprod(one_m, dim=1)
        Texcl = torch.cat([torch.ones_like(Tincl[:, :1]), Tincl[:, :-1]], dim=1)
        stop_here = valid & (Tincl.detach() < 1e-4)
        stopped = torch.cumsum(stop_here.to(torch.int32), dim=1) > 0
        live = valid & ~stopped
        w = torch.where(live, a * Texcl, torch.zeros_like(a))
        Cpix = w @ col                                              # [npix,3]
        Tfin = torch.prod(torch.where(live, one_m, torch.ones_like(one_m)), dim=1)
        idx = torch.arange(1, e - s + 1)[None, :].expand_as(live)
        last = torch.where(live, idx, torch.zeros_like(idx)).max(dim=1).values
        tile_img = Cpix.t() + Tfin[None, :] * bg.view(3, 1)
        pieces.append((y0, y1, x0, x1, tile_img, Tfin.detach(), last))
    if pieces:
        # assemble without in-place ops on a leaf-dependent tensor
        canvas = [[None] * gx for _ in range(gy)]
        for (y0, y1, x0, x1, img, Tf, last) in pieces:
            canvas[y0 // 16][x0 // 16] = img.view(3, y1 - y0, x1 - x0)
            final_T[y0:y1, x0:x1] = Tf.view(y1 - y0, x1 - x0)
            n_contrib[y0:y1, x0:x1] = last.view(y1 - y0, x1 - x0)
        rows = []
        for ty in range(gy):
            y0, y1 = ty * 16, min(ty * 16 + 16, H)
            row = []
            for tx in range(gx):
                x0, x1 = tx * 16, min(tx * 16 + 16, W)
                blk = canvas[ty][tx]
                if blk is None:
                    blk = bg.view(3, 1, 1).expand(3, y1 - y0, x1 - x0)
                row.append(blk)
            rows.append(torch.cat(row, dim=2))
        out = torch.cat(rows, dim=1)
    return out, final_T, n_contrib


def rasterize(means3D, means2D, opacities, *, shs=None, colors_precomp=None, scales=None, rotations=None,
              cov3D_precomp=None, H, W, tanfovx, tanfovy, bg, viewmatrix, projmatrix, campos,
              sh_degree=0, scale_modifier=1.0, return_aux=False):
    """Differentiable forward; mirrors GaussianRasterizer.forward + settings."""
    pre = preprocess(means3D, means2D, opacities, scales, rotations, cov3D_precomp, shs, colors_precomp,
                     H=H, W=W, tanfovx=tanfovx, tanfovy=tanfovy, viewmatrix=viewmatrix,
                     projmatrix=projmatrix, campos=campos, sh_degree=sh_degree,
                     scale_modifier=scale_modifier)
    pre["opacity"] = opacities.reshape(-1)
    keys, plist, ranges = bin_tiles(pre, H, W)
    img, final_T, n_contrib = render(pre, plist, ranges, bg, H, W)
    if return_aux:
        return img, pre["radius"], dict(pre=pre, keys=keys, point_list=plist, ranges=ranges,
                                        final_T=final_T, n_contrib=n_contrib)
    return img, pre["radius"]


def forward_backward(scene, cam, bg, dL_dout, threads: Optional[int] = None):
    """Convenience used by tests and the CPU baseline: numpy scene/camera in, numpy grads out."""
    import numpy as np
    if threads:
        torch.set_num_threads(threads)
    t = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a)).float()
    leaf = lambda a: None if a is None else t(a).clone().requires_grad_(True)
    means3D, scales, rots = leaf(scene.means3D), leaf(scene.scales), leaf(scene.rotations)
    opac = leaf(scene.opacities)
    colors, shs = leaf(scene.colors), leaf(scene.shs)
    means2D = torch.zeros(scene.P, 3, requires_grad=True)
    img, radii, aux = rasterize(
        means3D, means2D, opac, shs=shs, colors_precomp=colors, scales=scales, rotations=rots,
        H=cam.H, W=cam.W, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy, bg=t(bg), viewmatrix=t(cam.viewmatrix),
        projmatrix=t(cam.projmatrix), campos=t(cam.campos), sh_degree=scene.sh_degree, return_aux=True)
    (img * t(dL_dout)).sum().backward()
    g = lambda x: None if x is None else (torch.zeros_like(x) if x.grad is None else x.grad).numpy()
    grads = dict(dL_dmeans3D=g(means3D), dL_dmeans2D=g(means2D), dL_dscales=g(scales), dL_drots=g(rots),
                 dL_dopacity=g(opac), dL_dcolors=g(colors), dL_dsh=g(shs))
    return img.detach().numpy(), radii.numpy(), aux, grads
