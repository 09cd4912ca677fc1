"""Sim -> render glue: mirror of /root/reference/modules/tune/utils.py
    diff_rasterization            :323-421
    compute_bindings_xyz          :424-448
    compute_bindings_F            :451-472
    preprocess_for_rasterization  :475-523   (multi-object scenes, experiments/inference.py:326)
and gaussian_binding_with_clip_v1 of modules/d3gs/utils/binding_utils.py:199-285
and the two pixel losses of modules/d3gs/utils/loss_utils.py:17-24.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import Tensor

from . import ops
from .rasterizer import get_rasterizer


class Bindings:
    """The K x Np particle->kernel weight matrix (<= 10 uniform entries per row,
    modules/d3gs/utils/binding_utils.py:265-274) in CSR form on the device.
    Accepts the reference's torch.sparse_coo_tensor (finetune.py:587-591)."""

    def __init__(self, rowptr: Tensor, col: Tensor, w: Tensor, n_particles: int):
        self.rowptr, self.col, self.w = rowptr.int().contiguous(), col.int().contiguous(), w.float().contiguous()
        self.n_kernels = self.rowptr.shape[0] - 1
        self.n_particles = int(n_particles)

    @classmethod
    def from_sparse(cls, B: Tensor) -> "Bindings":
        B = B.coalesce()
        K, Np = B.shape
        rows, cols = B.indices()
        counts = torch.bincount(rows, minlength=K)
        rowptr = torch.zeros(K + 1, dtype=torch.int32, device=B.device)
        rowptr[1:] = torch.cumsum(counts, 0)
        return cls(rowptr, cols, B.values(), Np)

    @property
    def shape(self):
        return (self.n_kernels, self.n_particles)


class _BindSpMM(torch.autograd.Function):
    @staticmethod
    def forward(ctx, dense: Tensor, b: Bindings):
        ctx.b = b
        return ops.bind_spmm(b.rowptr, b.col, b.w, dense.detach().contiguous(), b.n_particles)

    @staticmethod
    def backward(ctx, g: Tensor):
        b = ctx.b
        return ops.bind_spmm(b.rowptr, b.col, b.w, g.contiguous(), b.n_particles, transpose=True), None


def _as_bindings(bindings) -> Bindings:
    if isinstance(bindings, Bindings):
        return bindings
    cached = getattr(bindings, "_neuma_csr", None)
    if cached is None:
        cached = Bindings.from_sparse(bindings)
        try:
            bindings._neuma_csr = cached
        except Exception:
            pass
    return cached


def compute_bindings_xyz(p_curr: Tensor, p_prev: Tensor, k_prev: Tensor, bindings) -> Tensor:
    """tune/utils.py:424-448: k_curr = k_prev.detach() + B (p_curr - p_prev.detach())."""
    delta_x = p_curr - p_prev.detach()
    return k_prev.detach() + _BindSpMM.apply(delta_x, _as_bindings(bindings))


def compute_bindings_F(deform_grad: Tensor, bindings) -> Tensor:
    """tune/utils.py:451-472."""
    return _BindSpMM.apply(torch.reshape(deform_grad, (-1, 9)), _as_bindings(bindings)).reshape(-1, 3, 3)


def diff_rasterization(x: Tensor, deform_grad: Optional[Tensor], gaussians, view_cam, background_color: Tensor,
                       gaussians_active_sh: Optional[int] = None, guassians_cov: Optional[Tensor] = None,
                       gaussians_opa: Optional[Tensor] = None, gaussians_shs: Optional[Tensor] = None,
                       scaling_modifier: Optional[float] = 1., force_mask_data: Optional[bool] = False) -> Tensor:
    """tune/utils.py:323-421 (argument names kept, typos included)."""
    means3D = x
    if gaussians is not None:
        cov3D_precomp = gaussians.get_covariance(scaling_modifier=scaling_modifier)
        opacity = gaussians.get_opacity
        shs = gaussians.get_features
        sh_degree = gaussians.active_sh_degree
    else:
        cov3D_precomp, opacity, shs, sh_degree = guassians_cov, gaussians_opa, gaussians_shs, gaussians_active_sh
    assert means3D.shape[0] == cov3D_precomp.shape[0], \
        f"Shape mismatch: means3D {means3D.shape[0]} cov3D {cov3D_precomp.shape[0]}"
    if deform_grad is not None:
        tensor_F = torch.reshape(deform_grad, (-1, 3, 3))
        assert cov3D_precomp.shape[0] == tensor_F.shape[0], \
            f"Shape mismatch: cov3D {cov3D_precomp.shape[0]} F {tensor_F.shape[0]}"
        # forward-only in the reference too: a bare wp.launch outside any tape
        cov3D_deformed = ops.deform_cov(cov3D_precomp.detach().reshape(-1, 6), tensor_F.detach())
    else:
        cov3D_deformed = cov3D_precomp
    means2D = torch.zeros_like(means3D, dtype=means3D.dtype, requires_grad=True) + 0
    rasterizer = get_rasterizer(view_cam, sh_degree, debug=False, bg_color=background_color)
    if force_mask_data:
        rendered_image, _ = rasterizer(means3D=means3D, means2D=means2D, shs=None,
                                       colors_precomp=torch.ones(means3D.shape[0], 3, device=x.device),
                                       opacities=opacity, scales=None, rotations=None,
                                       cov3D_precomp=cov3D_deformed)
    else:
        rendered_image, _ = rasterizer(means3D=means3D, means2D=means2D, shs=shs, colors_precomp=None,
                                       opacities=opacity, scales=None, rotations=None,
                                       cov3D_precomp=cov3D_deformed)
    return rendered_image


def preprocess_for_rasterization(obj_gaussians, obj_deform_grad, obj_kernels_prev, obj_particles_curr,
                                 obj_particles_prev, obj_bindings, obj_scalings):
    """tune/utils.py:475-523: per object, move the kernels with their bound particles and average the
    particles' deformation gradients, then concatenate everything diff_rasterization needs
    (pass the result as x=pack["means3D"], deform_grad=pack["deform_grad"], gaussians=None,
    guassians_cov=pack["cov3D"], ..., as experiments/inference.py:326-345 does)."""
    obj_x = [compute_bindings_xyz(p_curr, p_prev, k_prev, b)
             for p_curr, p_prev, k_prev, b in zip(obj_particles_curr, obj_particles_prev, obj_kernels_prev, obj_bindings)]
    obj_F = [compute_bindings_F(dg, b) for dg, b in zip(obj_deform_grad, obj_bindings)]
    obj_cov, obj_opa, obj_shs = [], [], []
    for gaussians, scaling in zip(obj_gaussians, obj_scalings):
        obj_cov.append(gaussians.get_covariance(scaling_modifier=scaling))
        obj_opa.append(gaussians.get_opacity)
        obj_shs.append(gaussians.get_features)
    return {
        "means3D": torch.cat(obj_x, dim=0),
        "deform_grad": torch.cat(obj_F, dim=0),
        "cov3D": torch.cat(obj_cov, dim=0),
        "opacity": torch.cat(obj_opa, dim=0),
        "shs": torch.cat(obj_shs, dim=0),
        "active_sh_degree": obj_gaussians[0].active_sh_degree,
    }


def gaussian_binding_with_clip_v1(gaussians, particles: Tensor, confidence: float = 0.95, max_particles: int = 10):
    """modules/d3gs/utils/binding_utils.py:199-285 on the GPU in one launch (the reference loops over
    the K kernels in Python).  Returns the K x Np weight matrix as a torch.sparse_coo_tensor -- what
    the reference stores in bindings.pt (tune/utils.py:297-318) -- instead of its dense intermediate
    (20 GB at 50k x 100k).  Uniform weights 1/n over the <= max_particles nearest particles inside the
    chi2_3(confidence) ellipsoid (binding_utils.py:265-274)."""
    from scipy.stats import chi2
    means = gaussians.get_xyz.detach().float().contiguous()
    cov6 = gaussians.get_covariance().detach().float().reshape(-1, 6).contiguous()
    x = particles.detach().float().contiguous()
    counts, cols, _ = ops.bind_build(means, cov6, x, float(chi2.ppf(confidence, 3)), max_particles)
    n_sel = counts.clamp(max=max_particles)
    assert int(n_sel.min()) > 0, "a kernel has no particle inside its confidence ellipsoid (binding_utils.py:281)"
    keep = cols >= 0
    rows = torch.arange(means.shape[0], device=x.device).unsqueeze(1).expand_as(cols)[keep]
    w = (1.0 / n_sel.float()).unsqueeze(1).expand_as(cols)[keep]
    return torch.sparse_coo_tensor(torch.stack([rows, cols[keep].long()]), w, (means.shape[0], x.shape[0])).coalesce()


def l1_loss(network_output: Tensor, gt: Tensor) -> Tensor:
    """modules/d3gs/utils/loss_utils.py:17-18."""
    return torch.abs((network_output - gt)).mean()


def l2_loss(network_output: Tensor, gt: Tensor) -> Tensor:
    """modules/d3gs/utils/loss_utils.py:23-24."""
    return ((network_output - gt) ** 2).mean()
