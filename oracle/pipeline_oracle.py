"""CPU oracle of the coupled rollout -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Restates the hot loop of /root/reference/experiments/finetune.py:360-414 (and the
forward-only loop of experiments/render.py:304-332) on top of the per-stage oracles:

    stress = elasticity(F); x,v,C,F = sim(...); F = plasticity(F)        finetune.py:362-364
    de_x = (x - center) / size                                          nclaw/utils.py:110-118
    means3D = k_prev + B (de_x - de_x_prev); F_k = B vec(F)             tune/utils.py:424-472
    cov3D' = F_k cov3D F_k^T  (NO gradient)                 tune/utils.py:353-373
    image = rasterizer(...); loss += decay * l2(image, gt)              finetune.py:377-389

Gradients are torch autograd over this forward (fp64 in the tests).
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch

from . import constitutive_oracle as CO
from . import mpm_oracle as MO
from . import raster_oracle as RO


def rollout_loss(c: MO.MPMConst, x, v, C, F, vol, rho, clip, enabled,
                 We, lora_e, Wp, lora_p, alpha: float,
                 B_sparse, k0, cov6, opac, shs, settings: RO.Settings, gts: Sequence[torch.Tensor],
                 substeps: int, center, size, dtype=torch.float64, loss="l2", colors_precomp=None,
                 return_images=False, weights=None):
    """One epoch of finetune_constitutive over len(gts) frames; returns the scalar loss.
    `settings` may be a list of cameras (`for view in used_views`, finetune.py:377): gts[f] is then a
    list with one image per view.  gts[f] = None: the frame is in `exclude_steps` -- `continue`
    BEFORE the binding reference is advanced (finetune.py:369-371).  weights[f] = the frame's
    rollout_decay_rate (finetune.py:388-389)."""
    cast = lambda t: None if t is None else t.to(dtype)
    x, v, C, F = map(cast, (x, v, C, F))
    vol, rho, clip = map(cast, (vol, rho, clip))
    center, size = cast(center), cast(size)
    de_x_prev = ((x - center) / size).detach()
    k_prev = cast(k0).detach()
    total = torch.zeros((), dtype=dtype)
    images = []
    cams = settings if isinstance(settings, (list, tuple)) else [settings]
    for fi, gt in enumerate(gts):
        for _ in range(substeps):
            stress = CO.elasticity(F, We, lora_e)
            x, v, C, F = MO.substep_fwd(c, x, v, C, F, stress, vol, rho, clip, enabled)
            F = CO.plasticity(F, Wp, alpha, lora_p)
        if gt is None:
            continue
        w = 1.0 if weights is None else float(weights[fi])
        de_x = (x - center) / size
        means3D = RO.bindings_xyz(de_x, de_x_prev, k_prev, B_sparse)
        Fk = RO.bindings_F(F, B_sparse)
        cov_def = RO.deform_cov(cast(cov6), Fk.detach()).detach()
        gt_views = gt if isinstance(gt, (list, tuple)) else [gt]
        for cam, gtv in zip(cams, gt_views):
            binning = RO.bin_np(cam, means3D.detach().float().numpy(), cov_def.float().numpy())
            img = RO.render_torch(cam, binning, means3D, None if colors_precomp is not None else cast(shs),
                                  cast(colors_precomp), cast(opac), cov_def, dtype)
            diff = img - cast(gtv)
            total = total + w * ((diff ** 2).mean() if loss == "l2" else diff.abs().mean())
            images.append(img.detach())
        de_x_prev = de_x.detach()
        k_prev = means3D.detach()
    if return_images:
        return total, images, (x, v, C, F)
    return total
