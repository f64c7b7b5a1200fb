"""Fused attribute head: CUDA kernels through the C ABI against the PyTorch
restatement of the reference's GSLayer.forward + attribute blending (oracle/attr_ref.py)."""
import numpy as np
import pytest
import torch

from oracle.attr_ref import activate_and_blend_ref, _TruncExp


def _inputs(P, seed, device="cpu", blend=True):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g)
    d = dict(xyz_raw=r(P, 3) * 0.01, pts=r(P, 3) * 0.1, scaling_raw=r(P, 3) * 0.5 - 5.0, rotation_raw=r(P, 4),
             opacity_raw=r(P, 1), rgb_raw=r(P, 3))
    d["scaling_raw"][0] = 16.0          # beyond the trunc_exp clamp: backward uses exp(15)
    d["rotation_raw"][1] = 0.0          # zero quaternion: the eps branch of normalize
    if blend:
        d.update(xyz_b=r(P, 3) * 0.01, opacity_b=r(P) * 0.1, color_w=1.0 + 0.1 * r(P, 48), color_b=0.1 * r(P, 48))
    return {k: v.to(device) for k, v in d.items()}


def test_oracle_matches_closed_forms():
    """The restatement against hand-written closed forms (CPU)."""
    d = _inputs(64, 0)
    m, s, q, o, c = activate_and_blend_ref(**d)
    assert torch.allclose(s, torch.exp(d["scaling_raw"]))
    n = d["rotation_raw"].norm(dim=1, keepdim=True).clamp_min(1e-12)
    assert torch.allclose(q, d["rotation_raw"] / n)
    assert torch.allclose(o, torch.sigmoid(d["opacity_raw"]) + d["opacity_b"].view(-1, 1))
    cw, cb = d["color_w"].view(-1, 16, 3), d["color_b"].view(-1, 16, 3)
    assert torch.allclose(c, torch.sigmoid(d["rgb_raw"]) * cw[:, 0] + cw[:, 1] - 1 + cb[:, 0])
    assert torch.allclose(m, d["xyz_raw"] + d["pts"] + d["xyz_b"])
    x = torch.tensor([1.0, 20.0], requires_grad=True)
    _TruncExp.apply(x).sum().backward()
    assert torch.allclose(x.grad, torch.exp(torch.tensor([1.0, 15.0])))


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", [dict(), dict(restrict_offset=True, clip_scaling=0.02), dict(xyz_offset=False),
                                 dict(blend=False)])
def test_cuda_attribute_head_matches_oracle(cuda_device, cfg):
    from guassianhand_b200.attributes import activate_and_blend
    cfg = dict(cfg)
    blend = cfg.pop("blend", True)
    P = 5001
    ref_in = {k: v.double().requires_grad_(True) for k, v in _inputs(P, 3, blend=blend).items()}
    gpu_in = {k: v.to(cuda_device).requires_grad_(True) for k, v in _inputs(P, 3, blend=blend).items()}
    ref = activate_and_blend_ref(**ref_in, **cfg)
    out = activate_and_blend(**gpu_in, **cfg)
    gen = torch.Generator().manual_seed(9)
    for a, b in zip(out, ref):
        assert a.shape == b.shape
        assert torch.allclose(a.detach().cpu().double(), b.detach(), rtol=2e-6, atol=1e-7)
    cot = [torch.randn(b.shape, generator=gen) for b in ref]
    sum((b * w.double()).sum() for b, w in zip(ref, cot)).backward()
    sum((a * w.to(cuda_device)).sum() for a, w in zip(out, cot)).backward()
    for k in ref_in:
        gr, gg = ref_in[k].grad, gpu_in[k].grad
        if gr is None:
            assert gg is None or float(gg.abs().max()) == 0.0, k
            continue
        gg = gg.detach().cpu().double()
        den = float(gr.abs().max()) or 1.0
        assert float((gg - gr).abs().max()) / den <= 1e-5, k


def test_oracle_sh_path_matches_reference_expression():
    """use_rgb = False (:201-204, :329-334): no sigmoid on the SH logits; color_w applied twice with color_b."""
    d = _inputs(32, 1)
    d.pop("rgb_raw")
    g = torch.Generator().manual_seed(4)
    shs_raw = torch.randn(32, 48, generator=g)
    cw, cb = d["color_w"].view(-1, 16, 3), d["color_b"].view(-1, 16, 3)
    *_, shs = activate_and_blend_ref(**d, shs_raw=shs_raw)
    assert shs.shape == (32, 16, 3)
    assert torch.allclose(shs, shs_raw.view(-1, 16, 3) * cw * cw + cb)
    d.pop("color_b")
    *_, shs = activate_and_blend_ref(**d, shs_raw=shs_raw)
    assert torch.allclose(shs, shs_raw.view(-1, 16, 3) * cw)
    d.pop("color_w")
    *_, shs = activate_and_blend_ref(**d, shs_raw=shs_raw)
    assert torch.equal(shs, shs_raw.view(-1, 16, 3))


@pytest.mark.gpu
@pytest.mark.parametrize("terms", ["w+b", "w", "none"])
def test_cuda_sh_path_matches_oracle(cuda_device, terms):
    from guassianhand_b200.attributes import activate_and_blend
    P = 3001
    base = _inputs(P, 7)
    base.pop("rgb_raw")
    if terms != "w+b":
        base.pop("color_b")
    if terms == "none":
        base.pop("color_w")
    shs_raw = torch.randn(P, 48, generator=torch.Generator().manual_seed(8))
    ref_in = {k: v.double().requires_grad_(True) for k, v in {**base, "shs_raw": shs_raw}.items()}
    gpu_in = {k: v.to(cuda_device).requires_grad_(True) for k, v in {**base, "shs_raw": shs_raw}.items()}
    ref = activate_and_blend_ref(**ref_in)
    out = activate_and_blend(**gpu_in)
    assert out[4].shape == (P, 16, 3)
    for a, b in zip(out, ref):
        assert torch.allclose(a.detach().cpu().double(), b.detach(), rtol=2e-6, atol=1e-7)
    gen = torch.Generator().manual_seed(9)
    cot = [torch.randn(b.shape, generator=gen) for b in ref]
    sum((b * w.double()).sum() for b, w in zip(ref, cot)).backward()
    sum((a * w.to(cuda_device)).sum() for a, w in zip(out, cot)).backward()
    for k in ref_in:
        gr, gg = ref_in[k].grad, gpu_in[k].grad
        assert gr is not None and gg is not None, k
        den = float(gr.abs().max()) or 1.0
        assert float((gg.detach().cpu().double() - gr).abs().max()) / den <= 1e-5, k
    with pytest.raises(AttributeError):                       # color_b without color_w, as the reference
        activate_and_blend(**{k: v for k, v in gpu_in.items() if k not in ("color_w", "color_b")},
                           color_b=torch.zeros(P, 48, device=cuda_device))


@pytest.mark.gpu
def test_attribute_head_feeds_the_rasterizer(cuda_device):
    """Head -> rasterize_views -> loss: gradients reach the raw head outputs and blending terms."""
    from guassianhand_b200 import api, scenes
    from guassianhand_b200.attributes import activate_and_blend
    import util
    P = 4000
    d = {k: v.to(cuda_device).requires_grad_(True) for k, v in _inputs(P, 5).items()}
    m, s, q, o, c = activate_and_blend(**d)
    views = util.gpu_views(scenes.fibonacci_cameras(2, 64, 80, seed=1), np.zeros(3, np.float32), cuda_device)
    img, _ = api.rasterize_views(m, o, views, colors_precomp=c, scales=s, rotations=q)
    img.sum().backward()
    for k in ("xyz_raw", "scaling_raw", "rotation_raw", "opacity_raw", "rgb_raw", "xyz_b", "opacity_b", "color_w"):
        assert d[k].grad is not None and torch.isfinite(d[k].grad).all(), k
    assert float(d["rgb_raw"].grad.abs().sum()) > 0
