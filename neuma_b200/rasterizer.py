"""Drop-in for the `diff_gaussian_rasterization` Python API the reference imports at
/root/reference/modules/d3gs/gaussian_renderer/__init__.py:3 and calls at
modules/tune/utils.py:385-419: GaussianRasterizationSettings, GaussianRasterizer,
plus get_rasterizer (gaussian_renderer/__init__.py:92-119).
"""
from __future__ import annotations

import math
from typing import NamedTuple, Optional

import torch
import torch.nn as nn
from torch import Tensor

from . import ops


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: Tensor
    scale_modifier: float
    viewmatrix: Tensor
    projmatrix: Tensor
    sh_degree: int
    campos: Tensor
    prefiltered: bool
    debug: bool


def _csettings(rs: GaussianRasterizationSettings):
    return ops.raster_settings(rs.image_height, rs.image_width, rs.tanfovx, rs.tanfovy, rs.bg,
                               rs.scale_modifier, rs.viewmatrix, rs.projmatrix, rs.sh_degree, rs.campos,
                               rs.prefiltered, rs.debug)


class _RasterizeGaussians(torch.autograd.Function):

    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, cov3Ds_precomp, cs):
        m3 = means3D.detach().contiguous()
        sh_ = None if sh is None else sh.detach().contiguous()
        cp = None if colors_precomp is None else colors_precomp.detach().contiguous()
        op = opacities.detach().contiguous()
        cov = cov3Ds_precomp.detach().contiguous()
        color, radii = ops.raster_fwd(cs, m3, sh_, cp, op, cov)
        # several forwards run before their backwards in a rollout: keep this call's state
        ctx.saved_state = ops.raster_save_state(m3) if any(ctx.needs_input_grad) else None
        ctx.cs = cs
        ctx.save_for_backward(m3, op, cov, radii, *( [sh_] if sh_ is not None else [cp]))
        ctx.has_sh = sh_ is not None
        ctx.mark_non_differentiable(radii)
        return color, radii

    @staticmethod
    def backward(ctx, grad_color, _grad_radii):
        m3, op, cov, radii, col = ctx.saved_tensors
        sh, cp = (col, None) if ctx.has_sh else (None, col)
        need = ctx.needs_input_grad
        want_color = need[2] if ctx.has_sh else need[3]
        g3, g2, gcov, gop, gcol = ops.raster_bwd(ctx.cs, m3, sh, cp, op, cov, radii, grad_color.contiguous(),
                                                 want_cov=need[5], want_opacity=need[4], want_color=want_color,
                                                 saved=ctx.saved_state)
        if gop is not None:
            gop = gop.reshape(op.shape)
        return (g3 if need[0] else None, g2 if need[1] else None,
                gcol if (ctx.has_sh and need[2]) else None,
                gcol if (not ctx.has_sh and need[3]) else None, gop, gcov, None)


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings: GaussianRasterizationSettings):
        super().__init__()
        self.raster_settings = raster_settings
        self._cs = _csettings(raster_settings)

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None,
                rotations=None, cov3D_precomp=None):
        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception("Please provide excatly one of either SHs or precomputed colors!")
        if ((scales is None or rotations is None) and cov3D_precomp is None) or \
                ((scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception("Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!")
        if cov3D_precomp is None:
            cov3D_precomp = covariance_from_scaling_rotation(scales, self.raster_settings.scale_modifier, rotations)
        if means2D is None:
            means2D = torch.zeros_like(means3D)
        return _RasterizeGaussians.apply(means3D, means2D, shs, colors_precomp, opacities, cov3D_precomp, self._cs)


def covariance_from_scaling_rotation(scaling: Tensor, scaling_modifier: float, rotation: Tensor) -> Tensor:
    """build_covariance_from_scaling_rotation (modules/d3gs/scene/gaussian_model.py:27-31) with
    build_rotation / strip_lowerdiag (modules/d3gs/utils/general_utils.py:93-139); packed order
    xx, xy, xz, yy, yz, zz.  Constant per run in NeuMA (Gaussians are frozen): plain torch."""
    q = rotation / rotation.norm(dim=1, keepdim=True)
    r, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = torch.stack([
        1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
        2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
        2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], -1).reshape(-1, 3, 3)
    L = R * (scaling_modifier * scaling)[:, None, :]
    S = L @ L.transpose(1, 2)
    return torch.stack([S[:, 0, 0], S[:, 0, 1], S[:, 0, 2], S[:, 1, 1], S[:, 1, 2], S[:, 2, 2]], -1)


def get_rasterizer(viewpoint_camera, active_sh_degree: int, debug, bg_color: Tensor, scaling_modifier=1.0):
    """gaussian_renderer/__init__.py:92-119."""
    tanfovx = math.tan(viewpoint_camera.FoVx * 0.5)
    tanfovy = math.tan(viewpoint_camera.FoVy * 0.5)
    raster_settings = GaussianRasterizationSettings(
        image_height=int(viewpoint_camera.image_height),
        image_width=int(viewpoint_camera.image_width),
        tanfovx=tanfovx, tanfovy=tanfovy, bg=bg_color, scale_modifier=scaling_modifier,
        viewmatrix=viewpoint_camera.world_view_transform,
        projmatrix=viewpoint_camera.full_proj_transform,
        sh_degree=active_sh_degree, campos=viewpoint_camera.camera_center,
        prefiltered=False, debug=debug)
    return GaussianRasterizer(raster_settings=raster_settings)
