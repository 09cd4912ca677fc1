"""CPU oracle for the Gaussian-splat rasterizer -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.

PARITY UNPINNED: the rasterizer the reference calls (modules/tune/utils.py:385-419,
settings built at modules/d3gs/gaussian_renderer/__init__.py:92-119) is the external
`diff_gaussian_rasterization` submodule of graphdeco-inria/gaussian-splatting@b17ded92
(README.md:56-60).  Its source is not under /root/reference and it cannot be
installed here, so this module restates the *published* 3DGS algorithm
[UPSTREAM-UNVERIFIED]; the only in-repo anchors are the Python twin of the SH
evaluation (modules/d3gs/utils/sh_utils.py:57-112, checked in
tests/test_raster_oracle.py) and the camera conventions
(modules/d3gs/scene/cameras.py:54-57, utils/graphics_utils.py:38-73).

Two layers:
  * `bin_np`        numpy fp32, operation-for-operation the order the CUDA preprocess
                    uses (compiled with -fmad=false), so depth bits, radii, tile
                    rectangles and the sorted (tile<<32 | depth) keys are BIT-EXACT
                    comparable.
  * `render_torch`  differentiable torch (fp64 in the tests) preprocess + per-tile
                    alpha blending over the instance lists from `bin_np`; its autograd
                    is the backward oracle.  Two published quirks of the CUDA backward
                    are reproduced on purpose: the min(0.99, .) clamp passes gradient,
                    and 1/det^2 is regularised as 1/(det^2 + 1e-7); the view-space
                    clamp passes no gradient when active.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

BX = BY = 16
SH_C0 = 0.28209479177387814
SH_C1 = 0.4886025119029199
SH_C2 = [1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396]
SH_C3 = [-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154,
         -0.4570457994644658, 1.445305721320277, -0.5900435899266435]


@dataclass
class Settings:
    """GaussianRasterizationSettings (gaussian_renderer/__init__.py:103-116)."""
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: np.ndarray          # [3]
    viewmatrix: np.ndarray  # [4,4] transposed / row-vector convention (cameras.py:54)
    projmatrix: np.ndarray  # [4,4] full_proj_transform (cameras.py:56)
    sh_degree: int
    campos: np.ndarray      # [3]
    scale_modifier: float = 1.0


def look_at_camera(W=800, H=800, fov_deg=40.0, eye=(0.5, 0.5, 2.5), znear=0.01, zfar=100.0):
    """Pinhole camera of SURVEY 8d: at `eye`, looking along -z, built with the
    reference's getWorld2View2 / getProjectionMatrix formulas
    (modules/d3gs/utils/graphics_utils.py:38-73, cameras.py:48-57)."""
    fov = math.radians(fov_deg)
    # camera looks down +z in view space; world -z maps to view +z, world y up -> view -y
    R = np.diag([1.0, -1.0, -1.0])            # camera-to-world rotation (columns = camera axes)
    eye = np.asarray(eye, dtype=np.float64)
    T = -R.T @ eye                            # w2c translation
    Rt = np.zeros((4, 4))
    Rt[:3, :3] = R.T
    Rt[:3, 3] = T
    Rt[3, 3] = 1.0
    world_view = np.float32(Rt).T             # .transpose(0, 1), cameras.py:54
    t = math.tan(fov / 2)
    P = np.zeros((4, 4), dtype=np.float32)
    top, right = t * znear, t * znear
    P[0, 0] = 2.0 * znear / (2 * right)
    P[1, 1] = 2.0 * znear / (2 * top)
    P[3, 2] = 1.0
    P[2, 2] = zfar / (zfar - znear)
    P[2, 3] = -(zfar * znear) / (zfar - znear)
    proj = P.T
    full = (world_view @ proj).astype(np.float32)
    campos = np.linalg.inv(world_view)[3, :3].astype(np.float32)
    return Settings(H, W, t, t, np.zeros(3, np.float32), world_view.astype(np.float32), full, 3, campos)


# ----------------------------------------------------------------------------- fp32 binning
def bin_np(s: Settings, means3D: np.ndarray, cov6: np.ndarray):
    """fp32, same operation order as neuma_b200/csrc/raster_pre.cu (-fmad=false)."""
    f = np.float32
    p = means3D.astype(f)
    K = p.shape[0]
    vm = s.viewmatrix.astype(f).reshape(-1)
    pm = s.projmatrix.astype(f).reshape(-1)
    W, H = s.image_width, s.image_height
    tanx, tany = f(s.tanfovx), f(s.tanfovy)
    fx = f(W) / (f(2.0) * tanx)
    fy = f(H) / (f(2.0) * tany)
    tiles_x, tiles_y = (W + BX - 1) // BX, (H + BY - 1) // BY
    x, y, z = p[:, 0], p[:, 1], p[:, 2]

    def x43(m, k):
        return ((m[k] * x + m[4 + k] * y) + m[8 + k] * z) + m[12 + k]

    t0, t1, t2 = x43(vm, 0), x43(vm, 1), x43(vm, 2)
    depth = t2.copy()
    front = depth > f(0.2)
    ph = [x43(pm, k) for k in range(4)]
    with np.errstate(all="ignore"):
        pw = f(1.0) / (ph[3] + f(0.0000001))
        pp0, pp1 = ph[0] * pw, ph[1] * pw
        limx, limy = f(1.3) * tanx, f(1.3) * tany
        txtz, tytz = t0 / t2, t1 / t2
        c0 = np.minimum(limx, np.maximum(-limx, txtz)) * t2
        c1 = np.minimum(limy, np.maximum(-limy, tytz)) * t2
        J00 = fx / t2
        J02 = -(fx * c0) / (t2 * t2)
        J11 = fy / t2
        J12 = -(fy * c1) / (t2 * t2)
        T = np.zeros((2, 3, K), f)
        for j in range(3):
            T[0, j] = J00 * vm[4 * j + 0] + J02 * vm[4 * j + 2]
            T[1, j] = J11 * vm[4 * j + 1] + J12 * vm[4 * j + 2]
        cv = cov6.astype(f)
        S = [[cv[:, 0], cv[:, 1], cv[:, 2]], [cv[:, 1], cv[:, 3], cv[:, 4]], [cv[:, 2], cv[:, 4], cv[:, 5]]]
        TS = [[(T[i, 0] * S[0][j] + T[i, 1] * S[1][j]) + T[i, 2] * S[2][j] for j in range(3)] for i in range(2)]
        a = ((TS[0][0] * T[0, 0] + TS[0][1] * T[0, 1]) + TS[0][2] * T[0, 2]) + f(0.3)
        b = (TS[0][0] * T[1, 0] + TS[0][1] * T[1, 1]) + TS[0][2] * T[1, 2]
        c = ((TS[1][0] * T[1, 0] + TS[1][1] * T[1, 1]) + TS[1][2] * T[1, 2]) + f(0.3)
        det = a * c - b * b
        mid = f(0.5) * (a + c)
        disc = np.sqrt(np.maximum(f(0.1), mid * mid - det))
        l1, l2 = mid + disc, mid - disc
        radius = np.ceil(f(3.0) * np.sqrt(np.maximum(l1, l2)))
        px = ((pp0 + f(1.0)) * f(W) - f(1.0)) * f(0.5)
        py = ((pp1 + f(1.0)) * f(H) - f(1.0)) * f(0.5)
        ok = front & (det != 0) & np.isfinite(radius)
        radius_i = np.where(ok, radius, 0).astype(np.int64)
        rf = radius_i.astype(f)

        def clampi(v, hi):
            v = np.where(np.isfinite(v), v, 0)
            return np.minimum(hi, np.maximum(0, np.trunc(v).astype(np.int64)))

        x0 = clampi((px - rf) / f(BX), tiles_x)
        y0 = clampi((py - rf) / f(BY), tiles_y)
        x1 = clampi((px + rf + f(BX - 1)) / f(BX), tiles_x)
        y1 = clampi((py + rf + f(BY - 1)) / f(BY), tiles_y)
    touched = np.where(ok, (x1 - x0) * (y1 - y0), 0)
    vis = touched > 0
    radii = np.where(vis, radius_i, 0).astype(np.int32)
    dbits = depth.view(np.uint32).astype(np.uint64)
    keys, ids = [], []
    for i in np.nonzero(vis)[0]:
        ys, xs = np.meshgrid(np.arange(y0[i], y1[i]), np.arange(x0[i], x1[i]), indexing="ij")
        tid = (ys * tiles_x + xs).reshape(-1).astype(np.uint64)
        keys.append((tid << np.uint64(32)) | dbits[i])
        ids.append(np.full(tid.shape[0], i, np.int32))
    if keys:
        keys = np.concatenate(keys)
        ids = np.concatenate(ids)
        order = np.argsort(keys, kind="stable")   # LSD radix sort is stable
        keys, ids = keys[order], ids[order]
    else:
        keys, ids = np.zeros(0, np.uint64), np.zeros(0, np.int32)
    ntiles = tiles_x * tiles_y
    tile_of = (keys >> np.uint64(32)).astype(np.int64)
    ranges = np.zeros((ntiles, 2), np.int32)
    if keys.shape[0]:
        starts = np.searchsorted(tile_of, np.arange(ntiles), side="left")
        ends = np.searchsorted(tile_of, np.arange(ntiles), side="right")
        has = ends > starts
        ranges[has, 0] = starts[has]
        ranges[has, 1] = ends[has]
    return dict(radii=radii, keys=keys, ids=ids, ranges=ranges, visible=vis, depth=depth, px=px, py=py,
                tiles_x=tiles_x, tiles_y=tiles_y)


# ----------------------------------------------------------------------------- SH
def sh_to_rgb(deg: int, sh: torch.Tensor, dirs: torch.Tensor):
    """computeColorFromSH: same polynomials as modules/d3gs/utils/sh_utils.py:57-112,
    sh [K, M, 3], dirs [K,3] unit; returns pre-clamp rgb + 0.5."""
    x, y, z = dirs[:, 0:1], dirs[:, 1:2], dirs[:, 2:3]
    r = SH_C0 * sh[:, 0]
    if deg > 0:
        r = r - SH_C1 * y * sh[:, 1] + SH_C1 * z * sh[:, 2] - SH_C1 * x * sh[:, 3]
        if deg > 1:
            xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
            r = (r + SH_C2[0] * xy * sh[:, 4] + SH_C2[1] * yz * sh[:, 5] +
                 SH_C2[2] * (2.0 * zz - xx - yy) * sh[:, 6] + SH_C2[3] * xz * sh[:, 7] +
                 SH_C2[4] * (xx - yy) * sh[:, 8])
            if deg > 2:
                r = (r + SH_C3[0] * y * (3.0 * xx - yy) * sh[:, 9] + SH_C3[1] * xy * z * sh[:, 10] +
                     SH_C3[2] * y * (4.0 * zz - xx - yy) * sh[:, 11] +
                     SH_C3[3] * z * (2.0 * zz - 3.0 * xx - 3.0 * yy) * sh[:, 12] +
                     SH_C3[4] * x * (4.0 * zz - xx - yy) * sh[:, 13] +
                     SH_C3[5] * z * (xx - yy) * sh[:, 14] + SH_C3[6] * x * (xx - 3.0 * yy) * sh[:, 15])
    return r + 0.5


# ----------------------------------------------------------------------------- differentiable render
class _ConicFromCov(torch.autograd.Function):
    """conic = inverse(cov2D); backward with the published 1/(det^2 + 1e-7)."""

    @staticmethod
    def forward(ctx, a, b, c):
        det = a * c - b * b
        ctx.save_for_backward(a, b, c, det)
        return c / det, -b / det, a / det

    @staticmethod
    def backward(ctx, gA, gB, gC):
        a, b, c, det = ctx.saved_tensors
        reg = det * det / (det * det + 1e-7)
        A, B, C = c / det, -b / det, a / det
        g01 = 0.5 * gB
        m00, m01 = A * gA + B * g01, A * g01 + B * gC
        m10, m11 = B * gA + C * g01, B * g01 + C * gC
        s00 = -(m00 * A + m01 * B) * reg
        s01 = -(m00 * B + m01 * C) * reg
        s10 = -(m10 * A + m11 * B) * reg
        s11 = -(m10 * B + m11 * C) * reg
        return s00, s01 + s10, s11


def preprocess_torch(s: Settings, means3D, shs, colors_precomp, opacities, cov6, dtype=torch.float64):
    t = lambda a: torch.as_tensor(np.asarray(a), dtype=dtype)
    vm, pm = t(s.viewmatrix).reshape(-1), t(s.projmatrix).reshape(-1)
    W, H = s.image_width, s.image_height
    fx, fy = W / (2.0 * s.tanfovx), H / (2.0 * s.tanfovy)
    x, y, z = means3D[:, 0], means3D[:, 1], means3D[:, 2]

    def x43(m, k):
        return m[k] * x + m[4 + k] * y + m[8 + k] * z + m[12 + k]

    t0, t1, t2 = x43(vm, 0), x43(vm, 1), x43(vm, 2)
    ph = [x43(pm, k) for k in range(4)]
    pw = 1.0 / (ph[3] + 0.0000001)
    px = ((ph[0] * pw + 1.0) * W - 1.0) * 0.5
    py = ((ph[1] * pw + 1.0) * H - 1.0) * 0.5
    limx, limy = 1.3 * s.tanfovx, 1.3 * s.tanfovy
    txtz, tytz = t0 / t2, t1 / t2
    inx = (txtz >= -limx) & (txtz <= limx)
    iny = (tytz >= -limy) & (tytz <= limy)
    c0 = torch.where(inx, t0, (txtz.clamp(-limx, limx) * t2).detach())
    c1 = torch.where(iny, t1, (tytz.clamp(-limy, limy) * t2).detach())
    J00, J02 = fx / t2, -(fx * c0) / (t2 * t2)
    J11, J12 = fy / t2, -(fy * c1) / (t2 * t2)
    T0 = torch.stack([J00 * vm[4 * j + 0] + J02 * vm[4 * j + 2] for j in range(3)], -1)
    T1 = torch.stack([J11 * vm[4 * j + 1] + J12 * vm[4 * j + 2] for j in range(3)], -1)
    S = torch.stack([cov6[:, 0], cov6[:, 1], cov6[:, 2], cov6[:, 1], cov6[:, 3], cov6[:, 4],
                     cov6[:, 2], cov6[:, 4], cov6[:, 5]], -1).reshape(-1, 3, 3)
    a = torch.einsum("ni,nij,nj->n", T0, S, T0) + 0.3
    b = torch.einsum("ni,nij,nj->n", T0, S, T1)
    c = torch.einsum("ni,nij,nj->n", T1, S, T1) + 0.3
    A, B, C = _ConicFromCov.apply(a, b, c)
    if colors_precomp is not None:
        rgb = colors_precomp
    else:
        d = means3D - t(s.campos)
        d = d / d.norm(dim=1, keepdim=True)
        rgb = torch.clamp_min(sh_to_rgb(s.sh_degree, shs, d), 0.0)
    return dict(px=px, py=py, A=A, B=B, C=C, rgb=rgb, opacity=opacities.reshape(-1))


def render_torch(s: Settings, binning, means3D, shs, colors_precomp, opacities, cov6, dtype=torch.float64,
                 return_aux=False):
    """Alpha blending over the per-tile instance lists of `binning` (from bin_np)."""
    pre = preprocess_torch(s, means3D, shs, colors_precomp, opacities, cov6, dtype)
    W, H = s.image_width, s.image_height
    tiles_x = binning["tiles_x"]
    ids_all = torch.from_numpy(binning["ids"].astype(np.int64))
    bg = torch.as_tensor(np.asarray(s.bg), dtype=dtype)
    img = bg.reshape(3, 1, 1).expand(3, H, W).clone()
    n_contrib = torch.zeros(H, W, dtype=torch.int64)
    final_T = torch.ones(H, W, dtype=dtype)
    borderline = torch.zeros(H, W, dtype=torch.bool)
    for tile, (st, en) in enumerate(binning["ranges"]):
        if en <= st:
            continue
        ty, tx = divmod(tile, tiles_x)
        ys = torch.arange(ty * BY, min((ty + 1) * BY, H))
        xs = torch.arange(tx * BX, min((tx + 1) * BX, W))
        gy, gx = torch.meshgrid(ys, xs, indexing="ij")
        pfx, pfy = gx.reshape(-1, 1).to(dtype), gy.reshape(-1, 1).to(dtype)
        ids = ids_all[st:en]
        dx = pre["px"][ids][None, :] - pfx
        dy = pre["py"][ids][None, :] - pfy
        A, B, C = pre["A"][ids][None, :], pre["B"][ids][None, :], pre["C"][ids][None, :]
        power = -0.5 * (A * dx * dx + C * dy * dy) - B * dx * dy
        raw = pre["opacity"][ids][None, :] * torch.exp(power)
        alpha = raw + (torch.clamp_max(raw, 0.99) - raw).detach()   # clamp passes gradient (published bwd)
        valid = (power <= 0) & (alpha >= 1.0 / 255.0)
        a_eff = torch.where(valid, alpha, torch.zeros_like(alpha))
        one_m = 1.0 - a_eff
        T_incl = torch.cumprod(one_m, dim=1)
        T_excl = torch.cat([torch.ones_like(T_incl[:, :1]), T_incl[:, :-1]], dim=1)
        stop = valid & (T_incl < 0.0001)
        stopped = torch.cumsum(stop.to(torch.int64), dim=1) > 0      # inclusive: the stopping one is not blended
        contrib = valid & ~stopped
        wgt = torch.where(contrib, a_eff * T_excl, torch.zeros_like(a_eff))
        col = wgt @ pre["rgb"][ids]                                   # [pix, 3]
        # transmittance after the last blended Gaussian
        any_stop = stopped[:, -1]
        first_stop = torch.argmax(stopped.to(torch.int64), dim=1)
        Tf = torch.where(any_stop, T_excl.gather(1, first_stop[:, None])[:, 0], T_incl[:, -1])
        idx1 = torch.arange(1, ids.shape[0] + 1)[None, :].expand_as(contrib)
        last = torch.where(contrib, idx1, torch.zeros_like(idx1)).max(dim=1).values
        out = col + Tf[:, None] * bg[None, :]
        img[:, gy.reshape(-1), gx.reshape(-1)] = out.t()
        n_contrib[gy.reshape(-1), gx.reshape(-1)] = last
        final_T[gy.reshape(-1), gx.reshape(-1)] = Tf.detach()
        if return_aux:
            with torch.no_grad():
                live = ~stopped | stop
                near = ((alpha - 1.0 / 255.0).abs() < 5e-6) | ((T_incl - 0.0001).abs() < 5e-7) | (power.abs() < 1e-6)
                borderline[gy.reshape(-1), gx.reshape(-1)] = (near & live).any(dim=1)
    if return_aux:
        return img, dict(n_contrib=n_contrib, final_T=final_T, borderline=borderline, pre=pre)
    return img


def rasterize(s: Settings, means3D, shs, colors_precomp, opacities, cov6, dtype=torch.float64, return_aux=False):
    """GaussianRasterizer.forward(means3D, means2D, shs, colors_precomp, opacities,
    cov3D_precomp) -> (image [3,H,W], radii [K]) as called at tune/utils.py:392-419."""
    binning = bin_np(s, means3D.detach().cpu().numpy(), cov6.detach().cpu().numpy())
    cast = lambda v: None if v is None else v.to(dtype)
    res = render_torch(s, binning, cast(means3D), cast(shs), cast(colors_precomp), cast(opacities), cast(cov6),
                       dtype, return_aux)
    if return_aux:
        img, aux = res
        aux["binning"] = binning
        return img, torch.from_numpy(binning["radii"]), aux
    return res, torch.from_numpy(binning["radii"])


# ----------------------------------------------------------------------------- bindings / covariance
def deform_cov(cov6: torch.Tensor, F: torch.Tensor) -> torch.Tensor:
    """deform_cov_by_F, modules/d3gs/utils/simulation_utils.py:25-48."""
    S = torch.stack([cov6[:, 0], cov6[:, 1], cov6[:, 2], cov6[:, 1], cov6[:, 3], cov6[:, 4],
                     cov6[:, 2], cov6[:, 4], cov6[:, 5]], -1).reshape(-1, 3, 3)
    c = F @ S @ F.transpose(1, 2)
    return torch.stack([c[:, 0, 0], c[:, 0, 1], c[:, 0, 2], c[:, 1, 1], c[:, 1, 2], c[:, 2, 2]], -1)


def bindings_xyz(p_curr, p_prev, k_prev, B):
    """compute_bindings_xyz, modules/tune/utils.py:424-448 (B: torch sparse [K,Np])."""
    return k_prev.detach() + torch.sparse.mm(B, p_curr - p_prev.detach())


def bindings_F(F, B):
    """compute_bindings_F, modules/tune/utils.py:451-472."""
    return torch.sparse.mm(B, F.reshape(-1, 9)).reshape(-1, 3, 3)
