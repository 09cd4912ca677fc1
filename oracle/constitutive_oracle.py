"""CPU oracle for the neural constitutive law -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.

torch-CPU restatement of
    batch_svd + sign fix                  /root/reference/modules/nclaw/warp/svd.py:59-96
    InvariantFullMetaElasticity.forward   modules/nclaw/material/meta.py:196-221
    InvariantFullMetaPlasticity.forward   modules/nclaw/material/meta.py:468-489
    MLPBlock (bias-free Linear + exact GELU)   meta.py:20-42, material/utils.py:16-17
    LinearLoRA.forward (unmerged)         modules/nclaw/material/loralib.py:216-224

wp.svd3 (warp-lang 0.6.1, not installable here) is replaced by torch.linalg.svd;
the sign fix of svd.py:76-92 is applied on top, which lands both on the same
convention (U, V rotations, third singular value signed).  Gradients are torch
autograd over this forward (fp64 in the tests); the SVD adjoint is therefore the
textbook one and is only compared where the singular values are separated.
`polar_newton` gives R without an SVD, for gradient checks at F ~ I.

PARITY PIN: tests/golden/material_*.npz were produced by importing the reference's
own InvariantFullMeta{Elasticity,Plasticity} / LinearLoRA classes
(tests/golden/make_golden_material.py, with only `self.svd` swapped for
torch.linalg.svd + the reference's sign fix) on the shipped base_models weights.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import torch


def svd_signfix(F: torch.Tensor):
    """batch_svd, svd.py:59-96: U, sigma, Vh with det U, det V forced to +1 by
    flipping the last column of U / V and the last singular value (both fixes
    can fire, flipping sigma[2] twice)."""
    U, s, Vh = torch.linalg.svd(F)
    V = Vh.transpose(-1, -2)
    du = torch.linalg.det(U) < 0
    dv = torch.linalg.det(V) < 0
    flip_u = torch.where(du, -1.0, 1.0).to(F.dtype)
    flip_v = torch.where(dv, -1.0, 1.0).to(F.dtype)
    one = torch.ones_like(flip_u)
    U = U * torch.stack([one, one, flip_u], -1)[:, None, :]
    V = V * torch.stack([one, one, flip_v], -1)[:, None, :]
    s = s * torch.stack([one, one, flip_u * flip_v], -1)
    return U, s, V.transpose(-1, -2)


def polar_newton(F: torch.Tensor, iters: int = 40) -> torch.Tensor:
    """Rotation factor of the polar decomposition by Higham's Newton iteration
    (differentiable and smooth at repeated singular values)."""
    R = F
    for _ in range(iters):
        R = 0.5 * (R + torch.linalg.inv(R).transpose(-1, -2))
    return R


def linear(x, W, lora: Optional[Tuple[torch.Tensor, torch.Tensor, float]] = None):
    """nn.Linear(bias=False) or LinearLoRA.forward (unmerged), loralib.py:216-224."""
    y = x @ W.t()
    if lora is not None:
        A, B, scaling = lora
        y = y + (x @ A.t() @ B.t()) * scaling
    return y


def gelu(x):
    return torch.nn.functional.gelu(x)  # nn.GELU() default: exact erf


def invariants(F, normalize_input=True, R_override=None):
    """meta.py:197-211."""
    U, sigma, Vh = svd_signfix(F)
    R = U @ Vh if R_override is None else R_override
    Ft = F.transpose(-1, -2)
    FtF = Ft @ F
    I = torch.eye(3, dtype=F.dtype)
    if normalize_input:
        I1 = sigma - 1.0
        I2 = (FtF - I).reshape(-1, 9)
        I3 = torch.linalg.det(F).unsqueeze(1) - 1.0
    else:
        I1, I2, I3 = sigma, FtF.reshape(-1, 9), torch.linalg.det(F).unsqueeze(1)
    return torch.cat([I1, I2, I3], dim=1), R, Ft


def mlp(feat, W: Sequence[torch.Tensor], lora=None):
    lo = lora if lora is not None else [None, None, None]
    x = gelu(linear(feat, W[0], lo[0]))
    x = gelu(linear(x, W[1], lo[1]))
    x = linear(x, W[2], lo[2])
    x = x.reshape(-1, 3, 3)
    return 0.5 * (x.transpose(-1, -2) + x)


def elasticity(F, W, lora=None, normalize_input=True, R_override=None):
    """InvariantFullMetaElasticity.forward, meta.py:196-221 -> Kirchhoff stress."""
    feat, R, Ft = invariants(F, normalize_input, R_override)
    x = mlp(feat, W, lora)
    P = R @ x
    return P @ Ft


def plasticity(F, W, alpha, lora=None, R_override=None):
    """InvariantFullMetaPlasticity.forward, meta.py:468-489."""
    feat, R, _ = invariants(F, True, R_override)
    x = mlp(feat, W, lora)
    return alpha * (R @ x) + F


def merge_lora(W, lora):
    """W + scaling * B A (LinearLoRA.train(False) merge, loralib.py:203-207)."""
    if lora is None:
        return list(W)
    return [w if l is None else w + l[2] * (l[1] @ l[0]) for w, l in zip(W, lora)]
