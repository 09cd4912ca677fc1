"""Synthetic workloads of the shapes BASELINE.json names: particle
clouds, Gaussian kernels, particle->kernel bindings and a pinhole camera.  Host-side
setup only -- nothing here is on the timed path.
"""
from __future__ import annotations

import math
from types import SimpleNamespace
from typing import Optional

import numpy as np
import torch

CHI2_3_095 = 7.814727903251179  # chi^2_3(0.95): the Mahalanobis gate of binding_utils.py

# dt: the shipped configs use 1e-3 / 5e-4 at num_grids 32 (finetune-jd.yaml:35-45).  The explicit
# MLS-MPM step is CFL-limited (dt ~ dx / wave speed): measured on the B200 with the shipped jelly
# weights, 128^3 goes NaN within 10 substeps at dt = 1e-3 and is stable at 2.5e-4, so the scaled-up
# grids scale dt with dx (SURVEY 8d's dt = 1e-3 at 128^3 does not survive contact with the GPU).
WORKLOADS = {
    # name: Np, G, box lo/hi per axis, K, image, dt
    "bouncyball-5k-32": dict(n=5_000, G=32, ball=(0.5, 0.6, 0.5, 0.2), K=0, image=0, dt=1e-3),
    "bouncyball-50k-64": dict(n=50_000, G=64, ball=(0.5, 0.6, 0.5, 0.2), K=25_000, image=800, dt=1e-3),
    "jellyduck-100k-128": dict(n=100_000, G=128, box=((0.3, 0.7),) * 3, K=50_000, image=800, dt=2.5e-4),
    "claycat-200k-128": dict(n=200_000, G=128, box=((0.25, 0.75),) * 3, K=100_000, image=800, dt=2.5e-4),
    "block-1m-256": dict(n=1_000_000, G=256, box=((0.2, 0.8), (0.2, 0.8), (0.1, 0.9)), K=0, image=0, dt=1.25e-4),
    # a quarter of block-1m-256 along z at the same density: on 2 GPUs each rank carries what a rank of the
    # 8-GPU run carries (125 k particles) -- the development proxy for the 8-GPU slab behaviour
    "block-250k-256": dict(n=250_000, G=256, box=((0.2, 0.8), (0.2, 0.8), (0.4, 0.6)), K=0, image=0, dt=1.25e-4),
}


def make_particles(spec: dict, seed: int = 42):
    g = torch.Generator().manual_seed(seed)
    n = spec["n"]
    if "ball" in spec:
        cx, cy, cz, r = spec["ball"]
        pts = []
        need = n
        while need > 0:
            c = 2 * torch.rand(int(need * 2.2) + 16, 3, generator=g) - 1
            c = c[(c * c).sum(1) <= 1.0][:need]
            pts.append(c)
            need -= c.shape[0]
        x = torch.cat(pts)[:n] * r + torch.tensor([cx, cy, cz])
        volume = 4.0 / 3.0 * math.pi * r ** 3
    else:
        lo = torch.tensor([b[0] for b in spec["box"]])
        hi = torch.tensor([b[1] for b in spec["box"]])
        x = lo + (hi - lo) * torch.rand(n, 3, generator=g)
        volume = float(torch.prod(hi - lo))
    dx = 1.0 / spec["G"]
    assert float(x.max()) < 1 - 1.5 * dx and float(x.min()) > 1.5 * dx   # latent OOB
    v = torch.tensor([0.0, -0.5, 0.0]).expand(n, 3).contiguous()
    return dict(x=x.float(), v=v.float(), C=torch.zeros(n, 3, 3), F=torch.eye(3).repeat(n, 1, 1),
                vol=torch.full((n,), volume / n), rho=torch.full((n,), 1000.0), clip=torch.full((n,), 0.1),
                enabled=torch.ones(n, dtype=torch.int32))


def quat_to_rot(q: torch.Tensor) -> torch.Tensor:
    q = q / q.norm(dim=1, keepdim=True)
    r, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    return torch.stack([
        1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
        2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
        2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], -1).reshape(-1, 3, 3)


def make_gaussians(x: torch.Tensor, K: int, G: int, seed: int = 43, max_bind: int = 10):
    """Kernels centred on every (n/K)-th particle; scales log-uniform in [0.5, 2] dx;
    bindings = <= 10 nearest particles inside the chi^2_3(0.95) ellipsoid, uniform
    weights (modules/d3gs/utils/binding_utils.py:199-285, :265-274)."""
    from scipy.spatial import cKDTree
    g = torch.Generator().manual_seed(seed)
    n = x.shape[0]
    stride = max(n // K, 1)
    centers = x[::stride][:K].clone()
    K = centers.shape[0]
    dx = 1.0 / G
    scale = dx * torch.exp(torch.empty(K, 3).uniform_(math.log(0.5), math.log(2.0), generator=g))
    R = quat_to_rot(torch.randn(K, 4, generator=g))
    L = R * scale[:, None, :]
    S = L @ L.transpose(1, 2)
    cov6 = torch.stack([S[:, 0, 0], S[:, 0, 1], S[:, 0, 2], S[:, 1, 1], S[:, 1, 2], S[:, 2, 2]], -1)
    opac = torch.empty(K, 1).uniform_(0.3, 0.95, generator=g)
    shs = 0.3 * torch.randn(K, 16, 3, generator=g)
    tree = cKDTree(x.numpy())
    _, nbr = tree.query(centers.numpy(), k=max_bind)
    nbr = torch.from_numpy(np.asarray(nbr, dtype=np.int64)).reshape(K, -1)
    d = x[nbr] - centers[:, None, :]                                   # [K, 10, 3]
    Sinv = torch.linalg.inv(S.double()).float()
    maha = torch.einsum("kni,kij,knj->kn", d, Sinv, d)
    inside = maha <= CHI2_3_095
    inside[:, 0] |= ~inside.any(dim=1)                                 # a kernel is never left unbound
    counts = inside.sum(1)
    rowptr = torch.zeros(K + 1, dtype=torch.int32)
    rowptr[1:] = torch.cumsum(counts, 0)
    col = nbr[inside].to(torch.int32)
    w = (1.0 / counts.float()).repeat_interleave(counts)
    return dict(xyz=centers.float(), cov6=cov6.float(), opac=opac.float(), shs=shs.float(),
                rowptr=rowptr, col=col, w=w.float(), K=K)


def make_camera(W: int = 800, H: int = 800, fov_deg: float = 40.0, eye=(0.5, 0.5, 2.5),
                znear: float = 0.01, zfar: float = 100.0):
    """Pinhole camera at `eye` looking down -z, assembled exactly as the reference's
    Camera does (modules/d3gs/scene/cameras.py:48-57 with
    utils/graphics_utils.py::getWorld2View2 / getProjectionMatrix)."""
    fov = math.radians(fov_deg)
    R = np.diag([1.0, -1.0, -1.0])
    T = -R.T @ np.asarray(eye, dtype=np.float64)
    Rt = np.zeros((4, 4))
    Rt[:3, :3] = R.transpose()
    Rt[:3, 3] = T
    Rt[3, 3] = 1.0
    Rt = np.linalg.inv(np.linalg.inv(Rt))      # getWorld2View2 with translate=0, scale=1
    world_view = torch.tensor(np.float32(Rt)).transpose(0, 1)
    t = math.tan(fov / 2)
    top, right = t * znear, t * znear
    P = torch.zeros(4, 4)
    P[0, 0] = 2.0 * znear / (2 * right)
    P[1, 1] = 2.0 * znear / (2 * top)
    P[3, 2] = 1.0
    P[2, 2] = zfar / (zfar - znear)
    P[2, 3] = -(zfar * znear) / (zfar - znear)
    proj = P.transpose(0, 1)
    full = (world_view.unsqueeze(0).bmm(proj.unsqueeze(0))).squeeze(0)
    center = world_view.inverse()[3, :3]
    return SimpleNamespace(FoVx=fov, FoVy=fov, image_height=H, image_width=W, znear=znear, zfar=zfar,
                           world_view_transform=world_view, full_proj_transform=full, camera_center=center)
