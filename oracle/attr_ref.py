"""CPU oracle of the attribute head (test infrastructure; never imported by the product path).

Plain PyTorch restatement, operation by operation, of
    GSLayer.forward          /root/reference/tgs/models/renderer_one_shot.py:191-214
    forward_single_view      /root/reference/tgs/models/renderer_one_shot.py:298-334 (use_rgb and SH paths)
    _TruncExp                /root/reference/tgs/utils/ops.py:37-53
Gradients come from autograd (with the reference's own trunc_exp backward).  Pinned by importing the
reference's `_TruncExp` source semantics: forward exp(x), backward g * exp(clamp(x, max=15)).
"""
import torch


class _TruncExp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return torch.exp(x)

    @staticmethod
    def backward(ctx, g):
        x = ctx.saved_tensors[0]
        return g * torch.exp(torch.clamp(x, max=15))


def activate_and_blend_ref(xyz_raw, pts, scaling_raw, rotation_raw, opacity_raw, rgb_raw=None, xyz_b=None, opacity_b=None,
                           color_w=None, color_b=None, xyz_offset=True, restrict_offset=False, clip_scaling=None,
                           shs_raw=None):
    # GSLayer.forward (:191-214)
    rotation = torch.nn.functional.normalize(rotation_raw)
    scaling = _TruncExp.apply(scaling_raw)
    if clip_scaling is not None:
        scaling = torch.clamp(scaling, min=0, max=clip_scaling)
    opacity = torch.sigmoid(opacity_raw.reshape(-1, 1))
    use_rgb = shs_raw is None
    v = torch.sigmoid(rgb_raw) if use_rgb else shs_raw        # :201-204: sigmoid only with cfg.use_rgb
    shs = torch.reshape(v, (v.shape[0], -1, 3))
    v = xyz_raw
    if restrict_offset:
        max_step = 1.2 / 32
        v = (torch.sigmoid(v) - 0.5) * max_step
    xyz = v + pts if xyz_offset else pts
    # forward_single_view (:298-334)
    means3D = xyz
    if xyz_b is not None:
        means3D = means3D + xyz_b
    if opacity_b is not None:
        opacity = opacity + opacity_b.view(-1, 1)
    if not use_rgb:
        # :329-334, as written (color_w is applied twice when color_b is given)
        if color_w is not None:
            shs = shs * color_w.view(-1, 16, 3)
        if color_b is not None:
            shs = shs * color_w.view(-1, 16, 3) + color_b.view(-1, 16, 3)
        return means3D, scaling, rotation, opacity, shs
    colors_precomp = shs.squeeze(1)
    if color_w is not None:
        colors_precomp = colors_precomp * color_w.view(-1, 16, 3)[:, 0, :] + color_w.view(-1, 16, 3)[:, 1, :] - 1
    if color_b is not None:
        colors_precomp = colors_precomp + color_b.view(-1, 16, 3)[:, 0, :]
    return means3D, scaling, rotation, opacity, colors_precomp
