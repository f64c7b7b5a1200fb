"""Fused Gaussian-attribute head: activations + interaction-aware blending in one CUDA launch each
way.

Mirrors what the reference computes with ~15 elementwise PyTorch kernels per view between its Linear
heads and the rasterizer call:

    GSLayer.forward          /root/reference/tgs/models/renderer_one_shot.py:191-214
    forward_single_view      /root/reference/tgs/models/renderer_one_shot.py:298-334 (use_rgb and SH paths)
    trunc_exp                /root/reference/tgs/utils/ops.py:37-53

All math runs in libghr.so (csrc/attributes.cu) through the C ABI (ghr_attributes_forward /
ghr_attributes_backward); there is no PyTorch fallback.  The five outputs are carved out of ONE
allocation in the layout the rasterizer takes, so they can be handed to
GaussianRasterizer.forward / rasterize_views as they are.
"""
import ctypes as C
from typing import Optional

import torch

from . import _native as N
from .api import _f32c, _raw_stream, _require_cuda

_IN = ("xyz_raw", "pts", "scaling_raw", "rotation_raw", "opacity_raw", "rgb_raw", "xyz_b", "opacity_b", "color_w0",
       "color_w1", "color_b0")
_OUT = (("means3D", 3), ("scales", 3), ("rotations", 4), ("opacities", 1), ("colors", 3))


def _fill(a, P, flags, clip, tensors):
    a.P, a.flags, a.clip_scaling = P, flags, float(clip)
    for name, t in zip(_IN, tensors):
        setattr(a, name, None if t is None else t.data_ptr())


class _ActivateAndBlend(torch.autograd.Function):
    @staticmethod
    def forward(ctx, flags, clip, *tensors):
        tensors = tuple(None if t is None else _f32c(t) for t in tensors)
        _require_cuda(*tensors)
        xyz_raw = tensors[0]
        P, dev = xyz_raw.shape[0], xyz_raw.device
        # one allocation: means3D 3P | scales 3P | rotations 4P | opacities P | colors 3P  (P % 4 == 0 keeps
        # the float4 rotation block 16-byte aligned; otherwise pad each block)
        out_spec = _OUT if tensors[5] is not None else _OUT[:4]   # SH path: no colours from this kernel
        sizes = [(P * k + 3) // 4 * 4 for _, k in out_spec]
        flat = torch.empty(sum(sizes), dtype=torch.float32, device=dev)
        outs, o = [], 0
        for (name, k), n in zip(out_spec, sizes):
            outs.append(flat[o:o + P * k].view(P, k))
            o += n
        a = N.GhrAttributeArgs()
        _fill(a, P, flags, clip, tensors)
        for (name, _), t in zip(out_spec, outs):
            setattr(a, name, t.data_ptr())
        N.check(N.lib().ghr_attributes_forward(C.byref(a), _raw_stream(dev)), "ghr_attributes_forward")
        ctx.flags, ctx.clip = flags, clip
        ctx.present = [t is not None for t in tensors]
        ctx.save_for_backward(*[t for t in tensors if t is not None])
        if len(outs) == 4:
            outs.append(None)
        return tuple(outs)

    @staticmethod
    def backward(ctx, g_means3D, g_scales, g_rot, g_opac, g_col):
        saved = iter(ctx.saved_tensors)
        tensors = [next(saved) if p else None for p in ctx.present]
        xyz_raw = tensors[0]
        P, dev = xyz_raw.shape[0], xyz_raw.device
        a = N.GhrAttributeArgs()
        _fill(a, P, ctx.flags, ctx.clip, tensors)
        g = N.GhrAttributeGrads()
        keep = []
        for name, t in (("dL_dmeans3D", g_means3D), ("dL_dscales", g_scales), ("dL_drotations", g_rot),
                        ("dL_dopacity", g_opac), ("dL_dcolors", g_col)):
            if t is not None:
                t = _f32c(t)
                keep.append(t)
                setattr(g, name, t.data_ptr())
        grads = []
        for idx, (name, t) in enumerate(zip(_IN, tensors)):
            # needs_input_grad is offset by the two non-tensor arguments (flags, clip)
            if t is not None and ctx.needs_input_grad[2 + idx]:
                d = torch.empty_like(t)
                setattr(g, "d_" + name, d.data_ptr())
                grads.append(d)
            else:
                grads.append(None)
        N.check(N.lib().ghr_attributes_backward(C.byref(a), C.byref(g), _raw_stream(dev)), "ghr_attributes_backward")
        return (None, None, *grads)


class _ShBlend(torch.autograd.Function):
    """SH path of the attribute blending (renderer_one_shot.py:329-334): shs * color_w, and with color_b
    (shs * color_w) * color_w + color_b -- the reference's double multiplication kept."""

    @staticmethod
    def forward(ctx, x, w, b):
        x, w, b = (None if t is None else _f32c(t) for t in (x, w, b))
        _require_cuda(x, w, b)
        out = torch.empty_like(x)
        N.check(N.lib().ghr_sh_blend_forward(x.numel(), x.data_ptr(), None if w is None else w.data_ptr(),
                                             None if b is None else b.data_ptr(), out.data_ptr(),
                                             _raw_stream(x.device)), "ghr_sh_blend_forward")
        ctx.has = (w is not None, b is not None)
        ctx.save_for_backward(*[t for t in (x, w, b) if t is not None])
        return out

    @staticmethod
    def backward(ctx, g):
        saved = list(ctx.saved_tensors)
        x = saved.pop(0)
        w = saved.pop(0) if ctx.has[0] else None
        b = saved.pop(0) if ctx.has[1] else None
        g = _f32c(g)
        need = ctx.needs_input_grad
        d = [torch.empty_like(x) if (need[i] and t is not None) else None for i, t in enumerate((x, w, b))]
        p = lambda t: None if t is None else t.data_ptr()
        N.check(N.lib().ghr_sh_blend_backward(x.numel(), x.data_ptr(), p(w), p(b), g.data_ptr(), p(d[0]), p(d[1]),
                                              p(d[2]), _raw_stream(x.device)), "ghr_sh_blend_backward")
        return tuple(d)


def activate_and_blend(xyz_raw, pts, scaling_raw, rotation_raw, opacity_raw, rgb_raw=None,
                       xyz_b: Optional[torch.Tensor] = None, opacity_b: Optional[torch.Tensor] = None,
                       color_w: Optional[torch.Tensor] = None, color_b: Optional[torch.Tensor] = None,
                       xyz_offset: bool = True, restrict_offset: bool = False,
                       clip_scaling: Optional[float] = None, shs_raw: Optional[torch.Tensor] = None):
    """Head outputs (after the Linear layers) -> rasterizer inputs.

    xyz_raw, pts, scaling_raw, rgb_raw: [P,3]; rotation_raw: [P,4]; opacity_raw: [P,1] or [P].
    Blending terms as the reference passes them (renderer_one_shot.py:259-268): xyz_b [P,3],
    opacity_b [P], color_w / color_b [P,48] or [P,16,3] (rows 0 and 1 of color_w and row 0 of color_b
    are used on the use_rgb path, :323-328).
    Returns (means3D [P,3], scales [P,3], rotations [P,4], opacities [P,1], colors [P,3]),
    differentiable w.r.t. every tensor argument.

    SH path (cfg.use_rgb false, :201-204 and :329-334): pass shs_raw [P,48] or [P,16,3] INSTEAD of rgb_raw; the
    fifth output is then shs [P,16,3] = shs_raw (* color_w) -- and, with color_b, (shs_raw * color_w) * color_w +
    color_b, the reference's expression as written.  color_b without color_w raises as the reference does."""
    if (rgb_raw is None) == (shs_raw is None):
        raise ValueError("activate_and_blend: provide exactly one of rgb_raw (use_rgb) / shs_raw (SH path)")
    flags = (N.GHR_ATTR_XYZ_OFFSET if xyz_offset else 0) | (N.GHR_ATTR_RESTRICT_OFFSET if restrict_offset else 0) | \
            (N.GHR_ATTR_CLIP_SCALING if clip_scaling is not None else 0)
    P = xyz_raw.shape[0]
    ob = None if opacity_b is None else opacity_b.reshape(P)
    if shs_raw is not None:
        if color_b is not None and color_w is None:
            raise AttributeError("'NoneType' object has no attribute 'view'")     # renderer_one_shot.py:334
        K = shs_raw.numel() // (3 * P) if P else 16
        shape = (P, K, 3)
        m, s, q, o, _ = _ActivateAndBlend.apply(flags, 0.0 if clip_scaling is None else float(clip_scaling), xyz_raw,
                                                pts, scaling_raw, rotation_raw, opacity_raw.reshape(P), None, xyz_b,
                                                ob, None, None, None)
        x = shs_raw.reshape(shape)
        if color_w is None:
            return m, s, q, o, x
        return m, s, q, o, _ShBlend.apply(x, color_w.reshape(shape), None if color_b is None else color_b.reshape(shape))
    w0 = w1 = b0 = None
    if color_w is not None:
        cw = color_w.reshape(P, 16, 3)
        w0, w1 = cw[:, 0, :], cw[:, 1, :]
    if color_b is not None:
        b0 = color_b.reshape(P, 16, 3)[:, 0, :]
    return _ActivateAndBlend.apply(flags, 0.0 if clip_scaling is None else float(clip_scaling), xyz_raw, pts,
                                   scaling_raw, rotation_raw, opacity_raw.reshape(P), rgb_raw, xyz_b, ob, w0, w1, b0)
