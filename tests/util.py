"""Seeded synthetic inputs shared by the tests (SURVEY 8d conventions)."""
import math

import torch


def make_cloud(n, G, seed=0, lo=0.3, hi=0.7, dtype=torch.float32, perturb_F=0.05, stress_scale=50.0,
               disabled_frac=0.0):
    g = torch.Generator().manual_seed(seed)
    x = lo + (hi - lo) * torch.rand(n, 3, generator=g)
    v = torch.tensor([0.0, -0.5, 0.0]) + 0.3 * torch.randn(n, 3, generator=g)
    C = 2.0 * torch.randn(n, 3, 3, generator=g)
    F = torch.eye(3).expand(n, 3, 3) + perturb_F * torch.randn(n, 3, 3, generator=g)
    stress = stress_scale * torch.randn(n, 3, 3, generator=g)
    vol = torch.full((n,), (hi - lo) ** 3 / max(n, 1))
    rho = torch.full((n,), 1000.0)
    clip = torch.full((n,), 0.1)
    enabled = torch.ones(n, dtype=torch.int32)
    if disabled_frac > 0:
        enabled[torch.rand(n, generator=g) < disabled_frac] = 0
    return dict(x=x.to(dtype), v=v.to(dtype), C=C.to(dtype), F=F.contiguous().to(dtype),
                stress=stress.to(dtype), vol=vol.to(dtype), rho=rho.to(dtype), clip=clip.to(dtype),
                enabled=enabled)


def rel_err(a, b):
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))
