"""Synthetic Gaussian scenes for the rasterizer tests (SURVEY 8d conventions)."""
import numpy as np
import torch

from oracle import raster_oracle as RO


def quat_to_rot(q):
    """restates modules/d3gs/utils/general_utils.py::build_rotation (:104-127)."""
    q = q / q.norm(dim=1, keepdim=True)
    r, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = torch.stack([
        1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
        2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
        2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], -1).reshape(-1, 3, 3)
    return R


def cov6_from_scale_rot(scale, quat):
    """build_covariance_from_scaling_rotation (gaussian_model.py:27-31) + strip_lowerdiag."""
    L = quat_to_rot(quat) @ torch.diag_embed(scale)
    S = L @ L.transpose(1, 2)
    return torch.stack([S[:, 0, 0], S[:, 0, 1], S[:, 0, 2], S[:, 1, 1], S[:, 1, 2], S[:, 2, 2]], -1)


def make_scene(K, W=128, H=128, seed=0, lo=0.3, hi=0.7, dx=1 / 64, sh_degree=3, bg=(0.0, 0.0, 0.0)):
    g = torch.Generator().manual_seed(seed)
    s = RO.look_at_camera(W=W, H=H)
    s.sh_degree = sh_degree
    s.bg = np.asarray(bg, np.float32)
    means = lo + (hi - lo) * torch.rand(K, 3, generator=g)
    scale = dx * torch.exp(torch.empty(K, 3).uniform_(np.log(0.5), np.log(2.0), generator=g))
    quat = torch.randn(K, 4, generator=g)
    cov6 = cov6_from_scale_rot(scale, quat)
    opac = torch.empty(K, 1).uniform_(0.3, 0.95, generator=g)
    shs = 0.3 * torch.randn(K, 16, 3, generator=g)
    return s, dict(means=means.float(), cov6=cov6.float(), opac=opac.float(), shs=shs.float())
