"""Thin torch-tensor front end of the C ABI (include/neuma_b200.h).

Every function here takes CUDA fp32 tensors, hands their device pointers and the
current torch stream to libneuma_b200.so, and returns torch tensors.  torch is
used for memory and streams only; there is no torch (or CPU) compute fallback.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from ._lib import Context, HaloSideC, MPMConstC, RasterSettingsC, NeumaError, ptr

BC_CODES = {"noslip": 0, "freeslip": 1}


def _ctx(t: torch.Tensor) -> Context:
    if not t.is_cuda:
        raise NeumaError("neuma_b200 ops need CUDA tensors (sm_100a); there is no CPU fallback")
    return Context.get(t.device.index if t.device.index is not None else torch.cuda.current_device())


def _stream(t: torch.Tensor) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _f32(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        raise NeumaError(f"expected float32, got {t.dtype}")
    return t.contiguous()


def mpm_const(num_grids: int, dt: float, bound: int, gravity: Sequence[float], eps: float, bc: str,
              z_off: int = 0, nz: Optional[int] = None, cap_nodes: int = 0) -> MPMConstC:
    """MPMConstant (mpm.py:158-167, dx/inv_dx as mpm.py:516-517).  cap_nodes > 0: the rollout grids
    are COMPACT (they hold the node bounding box only, cap_nodes nodes each)."""
    if bc not in BC_CODES:
        raise ValueError("invalid boundary condition: {}".format(bc))  # mpm.py:550
    c = MPMConstC()
    c.num_grids = int(num_grids)
    c.dt = float(dt)
    c.bound = int(bound)
    c.gravity[0], c.gravity[1], c.gravity[2] = [float(g) for g in gravity]
    c.dx = 1 / num_grids
    c.inv_dx = float(num_grids)
    c.eps = float(eps)
    c.bc = BC_CODES[bc]
    c.z_off = int(z_off)
    c.nz = int(num_grids if nz is None else nz)
    c.compact = 1 if cap_nodes > 0 else 0
    c.cap_nodes = int(cap_nodes)
    c.sel_mode, c.sel_lo_end, c.sel_up_start = 0, 0, 0
    return c


# --------------------------------------------------------------------- MPM
def mpm_grid_clear(c: MPMConstC, like: torch.Tensor) -> None:
    _ctx(like).call("neuma_mpm_grid_clear", C.byref(c), _stream(like))


def mpm_p2g(c: MPMConstC, x, v, Cm, stress, vol, rho, enabled) -> None:
    x, v, Cm, stress, vol, rho = map(_f32, (x, v, Cm, stress, vol, rho))
    _ctx(x).call("neuma_mpm_p2g", C.byref(c), x.shape[0], ptr(x), ptr(v), ptr(Cm), ptr(stress),
                 ptr(vol), ptr(rho), ptr(enabled), _stream(x))


def mpm_grid_op(c: MPMConstC, like: torch.Tensor) -> None:
    _ctx(like).call("neuma_mpm_grid_op", C.byref(c), _stream(like))


def mpm_g2p(c: MPMConstC, x, F, clip_bound, enabled, out=None, lazy_grid_op: bool = True):
    x, F, clip_bound = map(_f32, (x, F, clip_bound))
    n = x.shape[0]
    if out is None:
        out = (torch.empty_like(x), torch.empty_like(x), torch.empty_like(F), torch.empty_like(F))
    x2, v2, C2, F2 = out
    _ctx(x).call("neuma_mpm_g2p", C.byref(c), n, ptr(x), ptr(F), ptr(clip_bound), ptr(enabled),
                 ptr(x2), ptr(v2), ptr(C2), ptr(F2), int(lazy_grid_op), _stream(x))
    return x2, v2, C2, F2


def mpm_grid_export(c: MPMConstC, like: torch.Tensor, want_v: bool = False):
    G, nz = c.num_grids, c.nz
    mv = torch.empty(G, G, nz, 3, device=like.device, dtype=torch.float32)
    m = torch.empty(G, G, nz, device=like.device, dtype=torch.float32)
    v = torch.empty(G, G, nz, 3, device=like.device, dtype=torch.float32) if want_v else None
    _ctx(like).call("neuma_mpm_grid_export", C.byref(c), ptr(mv), ptr(m), ptr(v), _stream(like))
    return mv, m, v


def mpm_halo_pack(c: MPMConstC, z: int, nplanes: int, buf: torch.Tensor) -> None:
    _ctx(buf).call("neuma_mpm_halo_pack", C.byref(c), int(z), int(nplanes), ptr(buf), _stream(buf))


def mpm_halo_unpack_add(c: MPMConstC, z: int, nplanes: int, buf: torch.Tensor) -> None:
    _ctx(buf).call("neuma_mpm_halo_unpack_add", C.byref(c), int(z), int(nplanes), ptr(buf), _stream(buf))


def mpm_substep_fwd(c: MPMConstC, x, v, Cm, F, stress, vol, rho, clip_bound, enabled, out=None):
    x, v, Cm, F, stress, vol, rho, clip_bound = map(_f32, (x, v, Cm, F, stress, vol, rho, clip_bound))
    if out is None:
        out = (torch.empty_like(x), torch.empty_like(v), torch.empty_like(Cm), torch.empty_like(F))
    x2, v2, C2, F2 = out
    _ctx(x).call("neuma_mpm_substep_fwd", C.byref(c), x.shape[0], ptr(x), ptr(v), ptr(Cm), ptr(F),
                 ptr(stress), ptr(vol), ptr(rho), ptr(clip_bound), ptr(enabled),
                 ptr(x2), ptr(v2), ptr(C2), ptr(F2), _stream(x))
    return x2, v2, C2, F2


def mpm_substep_bwd(c: MPMConstC, x, v, Cm, F, stress, vol, rho, clip_bound, enabled, v2,
                    gx2, gv2, gC2, gF2, need: Sequence[bool] = (True,) * 5):
    x, v, Cm, F, stress, vol, rho, clip_bound, v2 = map(
        _f32, (x, v, Cm, F, stress, vol, rho, clip_bound, v2))
    gx2, gv2, gC2, gF2 = map(_f32, (gx2, gv2, gC2, gF2))
    outs = [torch.empty_like(t) if nd else None for t, nd in zip((x, v, Cm, F, stress), need)]
    _ctx(x).call("neuma_mpm_substep_bwd", C.byref(c), x.shape[0], ptr(x), ptr(v), ptr(Cm), ptr(F),
                 ptr(stress), ptr(vol), ptr(rho), ptr(clip_bound), ptr(enabled), ptr(v2),
                 ptr(gx2), ptr(gv2), ptr(gC2), ptr(gF2), *[ptr(o) for o in outs], _stream(x))
    return tuple(outs)


# --------------------------------------------------------------------- SVD / constitutive
def svd3_fwd(A: torch.Tensor):
    A = _f32(A)
    n = A.shape[0]
    U = torch.empty_like(A)
    s = torch.empty(n, 3, device=A.device, dtype=torch.float32)
    Vh = torch.empty_like(A)
    _ctx(A).call("neuma_svd3_fwd", n, ptr(A), ptr(U), ptr(s), ptr(Vh), _stream(A))
    return U, s, Vh


def svd3_bwd(U, s, Vh, gU, gs, gVh):
    U, s, Vh = map(_f32, (U, s, Vh))
    gU, gs, gVh = [None if g is None else _f32(g) for g in (gU, gs, gVh)]
    gA = torch.empty_like(U)
    _ctx(U).call("neuma_svd3_bwd", U.shape[0], ptr(U), ptr(s), ptr(Vh), ptr(gU), ptr(gs), ptr(gVh),
                 ptr(gA), _stream(U))
    return gA


KIND_ELASTIC, KIND_PLASTIC, KIND_ELASTIC_RAW = 0, 1, 2


def constitutive_fwd(kind: int, F, W0, W1, W2, alpha: float = 0.0):
    F, W0, W1, W2 = map(_f32, (F, W0, W1, W2))
    out = torch.empty_like(F)
    _ctx(F).call("neuma_constitutive_fwd", int(kind), F.shape[0], W1.shape[0], ptr(F), ptr(W0),
                 ptr(W1), ptr(W2), C.c_float(alpha), ptr(out), _stream(F))
    return out


def constitutive_bwd(kind: int, F, W0, W1, W2, alpha: float, gout, want_wgrad: bool = True):
    F, W0, W1, W2, gout = map(_f32, (F, W0, W1, W2, gout))
    gF = torch.empty_like(F)
    gW = [torch.zeros_like(w) for w in (W0, W1, W2)] if want_wgrad else [None, None, None]
    _ctx(F).call("neuma_constitutive_bwd", int(kind), F.shape[0], W1.shape[0], ptr(F), ptr(W0),
                 ptr(W1), ptr(W2), C.c_float(alpha), ptr(gout), ptr(gF), ptr(gW[0]), ptr(gW[1]),
                 ptr(gW[2]), _stream(F))
    return gF, gW[0], gW[1], gW[2]


# --------------------------------------------------------------------- bindings
def bind_spmm(rowptr, col, w, dense: torch.Tensor, n_particles: int, transpose: bool = False):
    dense = _f32(dense)
    K = rowptr.shape[0] - 1
    ncol = dense.shape[1]
    rows = n_particles if transpose else K
    out = torch.empty(rows, ncol, device=dense.device, dtype=torch.float32)
    _ctx(dense).call("neuma_bind_spmm", K, int(n_particles), ptr(rowptr), ptr(col), ptr(w), ncol,
                     ptr(dense), ptr(out), int(transpose), _stream(dense))
    return out


def deform_cov(cov6: torch.Tensor, F: torch.Tensor):
    cov6, F = _f32(cov6), _f32(F)
    out = torch.empty_like(cov6)
    _ctx(cov6).call("neuma_deform_cov", cov6.shape[0], ptr(cov6), ptr(F), ptr(out), _stream(cov6))
    return out


# --------------------------------------------------------------------- rasterizer
def raster_settings(image_height, image_width, tanfovx, tanfovy, bg, scale_modifier, viewmatrix,
                    projmatrix, sh_degree, campos, prefiltered=False, debug=False) -> RasterSettingsC:
    s = RasterSettingsC()
    s.image_height, s.image_width = int(image_height), int(image_width)
    s.tanfovx, s.tanfovy = float(tanfovx), float(tanfovy)
    def flat(a):
        if torch.is_tensor(a):
            a = a.detach().cpu().numpy()
        return [float(v) for v in np.asarray(a, dtype=np.float64).reshape(-1)]

    s.bg[0], s.bg[1], s.bg[2] = flat(bg)
    s.scale_modifier = float(scale_modifier)
    vm, pm = flat(viewmatrix), flat(projmatrix)
    for i in range(16):
        s.viewmatrix[i] = vm[i]
        s.projmatrix[i] = pm[i]
    s.sh_degree = int(sh_degree)
    s.campos[0], s.campos[1], s.campos[2] = flat(campos)
    s.prefiltered, s.debug = int(bool(prefiltered)), int(bool(debug))
    return s


def raster_fwd(s: RasterSettingsC, means3D, shs, colors_precomp, opacities, cov6):
    means3D, opacities, cov6 = map(_f32, (means3D, opacities, cov6))
    shs = None if shs is None else _f32(shs)
    colors_precomp = None if colors_precomp is None else _f32(colors_precomp)
    K = means3D.shape[0]
    M = 0 if shs is None else shs.shape[1]
    image = torch.empty(3, s.image_height, s.image_width, device=means3D.device, dtype=torch.float32)
    radii = torch.empty(K, device=means3D.device, dtype=torch.int32)
    _ctx(means3D).call("neuma_raster_fwd", C.byref(s), K, M, ptr(means3D), ptr(shs),
                       ptr(colors_precomp), ptr(opacities), ptr(cov6), ptr(image), ptr(radii),
                       _stream(means3D))
    return image, radii


def raster_save_state(like: torch.Tensor):
    """Copy the forward state the backward needs out of the ctx (see neuma_b200.h)."""
    ctx = _ctx(like)
    nbytes = C.c_size_t(0)
    ctx.call("neuma_raster_state_bytes", C.byref(nbytes))
    n = C.c_int(0)
    ctx.call("neuma_raster_num_rendered", _stream(like), C.byref(n))
    blob = torch.empty(nbytes.value, dtype=torch.uint8, device=like.device)
    ctx.call("neuma_raster_state_save", ptr(blob), _stream(like))
    return blob, n.value


def raster_bwd(s: RasterSettingsC, means3D, shs, colors_precomp, opacities, cov6, radii, dL_dimage,
               want_cov=False, want_opacity=False, want_color=False, saved=None):
    means3D, opacities, cov6, dL_dimage = map(_f32, (means3D, opacities, cov6, dL_dimage))
    shs = None if shs is None else _f32(shs)
    colors_precomp = None if colors_precomp is None else _f32(colors_precomp)
    K = means3D.shape[0]
    M = 0 if shs is None else shs.shape[1]
    dev = means3D.device
    g3 = torch.empty(K, 3, device=dev, dtype=torch.float32)
    g2 = torch.empty(K, 3, device=dev, dtype=torch.float32)
    gcov = torch.empty(K, 6, device=dev, dtype=torch.float32) if want_cov else None
    gop = torch.empty_like(opacities) if want_opacity else None
    gcol = None
    if want_color:
        gcol = torch.empty_like(shs if shs is not None else colors_precomp)
    _ctx(means3D).call("neuma_raster_bwd", C.byref(s), K, M, ptr(means3D), ptr(shs),
                       ptr(colors_precomp), ptr(opacities), ptr(cov6), ptr(radii), ptr(dL_dimage),
                       ptr(None if saved is None else saved[0]), 0 if saved is None else int(saved[1]),
                       ptr(g3), ptr(g2), ptr(gcov), ptr(gop), ptr(gcol), _stream(means3D))
    return g3, g2, gcov, gop, gcol


def raster_check(like: torch.Tensor) -> int:
    """Waits for the last forward's instance count and raises if a forward (eager or replayed from a
    CUDA graph) produced more tile instances than the buffers hold.  Returns the count."""
    ctx = _ctx(like)
    n = C.c_int(0)
    ctx.call("neuma_raster_num_rendered", _stream(like), C.byref(n))
    return n.value


def raster_binning(like: torch.Tensor, tiles: int):
    """Sorted instance keys / Gaussian ids / tile ranges of the last forward."""
    ctx = _ctx(like)
    n = C.c_int(0)
    ctx.call("neuma_raster_num_rendered", _stream(like), C.byref(n))
    L = n.value
    keys = torch.empty(L, device=like.device, dtype=torch.int64)
    ids = torch.empty(L, device=like.device, dtype=torch.int32)
    ranges = torch.empty(tiles, 2, device=like.device, dtype=torch.int32)
    ctx.call("neuma_raster_export_binning", ptr(keys), ptr(ids), ptr(ranges), _stream(like))
    return keys, ids, ranges


# --------------------------------------------------------------------- allocation-free variants
# (static buffers: used by neuma_b200.pipeline under CUDA-graph capture)
def constitutive_fwd_into(kind: int, F, W0, W1, W2, alpha: float, out) -> None:
    _ctx(F).call("neuma_constitutive_fwd", int(kind), F.shape[0], W1.shape[0], ptr(F), ptr(W0),
                 ptr(W1), ptr(W2), C.c_float(alpha), ptr(out), _stream(F))


def constitutive_bwd_into(kind: int, F, W0, W1, W2, alpha: float, gout, gF, gW0, gW1, gW2) -> None:
    """gW* are accumulated into (not zeroed)."""
    _ctx(F).call("neuma_constitutive_bwd", int(kind), F.shape[0], W1.shape[0], ptr(F), ptr(W0),
                 ptr(W1), ptr(W2), C.c_float(alpha), ptr(gout), ptr(gF), ptr(gW0), ptr(gW1), ptr(gW2),
                 _stream(F))


def mpm_substep_bwd_into(c: MPMConstC, x, v, Cm, F, stress, vol, rho, clip_bound, enabled, v2,
                         gx2, gv2, gC2, gF2, gx, gv, gC, gF, gstress) -> None:
    _ctx(x).call("neuma_mpm_substep_bwd", C.byref(c), x.shape[0], ptr(x), ptr(v), ptr(Cm), ptr(F),
                 ptr(stress), ptr(vol), ptr(rho), ptr(clip_bound), ptr(enabled), ptr(v2),
                 ptr(gx2), ptr(gv2), ptr(gC2), ptr(gF2), ptr(gx), ptr(gv), ptr(gC), ptr(gF), ptr(gstress),
                 _stream(x))


# --------------------------------------------------------------------- MPM rollout path
# (caller-owned per-substep grids + device-side bounding boxes; see include/neuma_b200.h)
def mpm_box_from_x(c: MPMConstC, x, enabled, box) -> None:
    _ctx(x).call("neuma_mpm_box_from_x", C.byref(c), x.shape[0], ptr(x), ptr(enabled), ptr(box), _stream(x))


def mpm_box_clear(c: MPMConstC, box, grid0=None, grid1=None, box_reset=None) -> None:
    _ctx(box).call("neuma_mpm_box_clear", C.byref(c), ptr(box), ptr(grid0), ptr(grid1), ptr(box_reset),
                   _stream(box))


def mpm_p2g_to(c: MPMConstC, x, v, Cm, stress, vol, rho, enabled, grid, box=None) -> None:
    _ctx(x).call("neuma_mpm_p2g_to", C.byref(c), x.shape[0], ptr(x), ptr(v), ptr(Cm), ptr(stress),
                 ptr(vol), ptr(rho), ptr(enabled), ptr(box), ptr(grid), _stream(x))


def mpm_grid_op_box(c: MPMConstC, box, grid, adj_to_clear=None, next_grid=None, box_reset=None) -> None:
    """grid_op over the box; next_grid / box_reset: also prepare the next substep's slot
    (neuma_mpm_grid_op_next)."""
    _ctx(grid).call("neuma_mpm_grid_op_next", C.byref(c), ptr(box), ptr(grid), ptr(adj_to_clear), ptr(next_grid),
                    ptr(box_reset), _stream(grid))


def mpm_g2p_from(c: MPMConstC, x, F, clip_bound, enabled, grid_v, out, box_next=None, box=None) -> None:
    x2, v2, C2, F2 = out
    _ctx(x).call("neuma_mpm_g2p_from", C.byref(c), x.shape[0], ptr(x), ptr(F), ptr(clip_bound), ptr(enabled),
                 ptr(box), ptr(grid_v), ptr(x2), ptr(v2), ptr(C2), ptr(F2), ptr(box_next), _stream(x))


def mpm_rollout_fwd(c: MPMConstC, x, v, Cm, F, stress, vol, rho, clip_bound, enabled, out, grid, adj, box,
                    box_next) -> None:
    x2, v2, C2, F2 = out
    _ctx(x).call("neuma_mpm_rollout_fwd", C.byref(c), x.shape[0], ptr(x), ptr(v), ptr(Cm), ptr(F), ptr(stress),
                 ptr(vol), ptr(rho), ptr(clip_bound), ptr(enabled), ptr(x2), ptr(v2), ptr(C2), ptr(F2),
                 ptr(grid), ptr(adj), ptr(box), ptr(box_next), _stream(x))


def mpm_rollout_bwd(c: MPMConstC, x, v, Cm, F, stress, vol, rho, clip_bound, enabled, v2, C2, gx2, gv2, gC2, gF2,
                    grid_v, adj, box, gx, gv, gC, gF, gstress) -> None:
    _ctx(x).call("neuma_mpm_rollout_bwd", C.byref(c), x.shape[0], ptr(x), ptr(v), ptr(Cm), ptr(F), ptr(stress),
                 ptr(vol), ptr(rho), ptr(clip_bound), ptr(enabled), ptr(v2), ptr(C2), ptr(gx2), ptr(gv2),
                 ptr(gC2), ptr(gF2), ptr(grid_v), ptr(adj), ptr(box), ptr(gx), ptr(gv), ptr(gC), ptr(gF),
                 ptr(gstress), _stream(x))


def mpm_cell_order(c: MPMConstC, x: torch.Tensor) -> torch.Tensor:
    """perm (int32[n]): particle indices in cell (Morton) order of their stencil base."""
    x = _f32(x)
    perm = torch.empty(x.shape[0], device=x.device, dtype=torch.int32)
    _ctx(x).call("neuma_mpm_cell_order", C.byref(c), x.shape[0], ptr(x), ptr(perm), _stream(x))
    return perm


def gather_rows(src: torch.Tensor, idx: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[i] = src[idx[i]] for contiguous 32-bit rows (fp32 / int32)."""
    if src.element_size() != 4 or not src.is_contiguous():
        raise NeumaError("gather_rows: contiguous 32-bit rows expected")
    n = idx.shape[0]
    if out is None:
        out = torch.empty((n,) + tuple(src.shape[1:]), device=src.device, dtype=src.dtype)
    words = int(src.numel() // max(src.shape[0], 1))
    _ctx(src).call("neuma_gather_rows", n, words, ptr(src), ptr(idx), ptr(out), _stream(src))
    return out


def invert_perm(perm: torch.Tensor) -> torch.Tensor:
    inv = torch.empty_like(perm)
    _ctx(perm).call("neuma_invert_perm", perm.shape[0], ptr(perm), ptr(inv), _stream(perm))
    return inv


# --------------------------------------------------------------------- staged constitutive weights
def constitutive_stage(W0, W1, W2, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Split / lay out the three Linear weights once (tensor-core operand image); reuse the
    returned buffer for every constitutive_*_staged call until the weights change."""
    W0, W1, W2 = map(_f32, (W0, W1, W2))
    ctx = _ctx(W0)
    if out is None:
        out = torch.empty(int(ctx.lib.neuma_constitutive_staged_bytes()), dtype=torch.uint8, device=W0.device)
    ctx.call("neuma_constitutive_stage", W1.shape[0], ptr(W0), ptr(W1), ptr(W2), ptr(out), _stream(W0))
    return out


def constitutive_fwd_staged(kind: int, F, staged, alpha: float, out, hidden: int = 64) -> None:
    _ctx(F).call("neuma_constitutive_fwd_staged", int(kind), F.shape[0], hidden, ptr(F), ptr(staged),
                 C.c_float(alpha), ptr(out), _stream(F))


def constitutive_bwd_staged(kind: int, F, staged, alpha: float, gout, gF, gW0, gW1, gW2, hidden: int = 64) -> None:
    """gW* are accumulated into (not zeroed)."""
    _ctx(F).call("neuma_constitutive_bwd_staged", int(kind), F.shape[0], hidden, ptr(F), ptr(staged),
                 C.c_float(alpha), ptr(gout), ptr(gF), ptr(gW0), ptr(gW1), ptr(gW2), _stream(F))


def constitutive_partials(like: torch.Tensor) -> torch.Tensor:
    """Zeroed per-CTA weight-gradient partial-sum buffer for constitutive_bwd_partials."""
    ctx = _ctx(like)
    nfl = int(ctx.lib.neuma_constitutive_partials_floats(ctx.handle))
    return torch.zeros(nfl, dtype=torch.float32, device=like.device)


def constitutive_bwd_partials(kind: int, F, staged, alpha: float, gout, gF, gWpart, hidden: int = 64,
                              accumulate_gF: bool = False) -> None:
    _ctx(F).call("neuma_constitutive_bwd_partials", int(kind), F.shape[0], hidden, ptr(F), ptr(staged),
                 C.c_float(alpha), ptr(gout), ptr(gF), int(accumulate_gF), ptr(gWpart), _stream(F))


def constitutive_fwd_p2g(kind: int, F, staged, stress, c: MPMConstC, x, v, Cm, vol, rho, enabled, grid, box=None,
                         hidden: int = 64) -> None:
    """elasticity + P2G in one kernel (neuma_constitutive_fwd_p2g)."""
    _ctx(F).call("neuma_constitutive_fwd_p2g", int(kind), F.shape[0], hidden, ptr(F), ptr(staged), ptr(stress),
                 C.byref(c), ptr(x), ptr(v), ptr(Cm), ptr(vol), ptr(rho), ptr(enabled), ptr(box), ptr(grid),
                 _stream(F))


def constitutive_bwd_scatter(Ftrial, staged, alpha: float, gout, gFtrial, gWpart, c: MPMConstC, x, F, clip_bound,
                             enabled, v2, gx2, gv2, gC2, adj, box=None, hidden: int = 64) -> None:
    """plasticity backward + adjoint scatter of G2P in one kernel (neuma_constitutive_bwd_scatter)."""
    _ctx(F).call("neuma_constitutive_bwd_scatter", Ftrial.shape[0], hidden, ptr(Ftrial), ptr(staged),
                 C.c_float(alpha), ptr(gout), ptr(gFtrial), ptr(gWpart), C.byref(c), ptr(x), ptr(F),
                 ptr(clip_bound), ptr(enabled), ptr(v2), ptr(gx2), ptr(gv2), ptr(gC2), ptr(box), ptr(adj),
                 _stream(F))


def constitutive_fwd_g2p(staged, alpha: float, F_out, c: MPMConstC, x, F, clip_bound, enabled, grid_v, out,
                         box_next=None, box=None, hidden: int = 64) -> None:
    """G2P + plasticity in one kernel (neuma_constitutive_fwd_g2p): out = (x', v', C', F_trial) as
    mpm_g2p_from writes them, F_out = plasticity(F_trial)."""
    x2, v2, C2, Ft = out
    _ctx(x).call("neuma_constitutive_fwd_g2p", x.shape[0], hidden, ptr(staged), C.c_float(alpha), ptr(F_out),
                 C.byref(c), ptr(x), ptr(F), ptr(clip_bound), ptr(enabled), ptr(box), ptr(grid_v), ptr(x2), ptr(v2),
                 ptr(C2), ptr(Ft), ptr(box_next), _stream(x))


def constitutive_bwd_gather(kind: int, staged, gWpart, c: MPMConstC, x, v, Cm, F, stress, vol, rho, clip_bound,
                            enabled, v2, C2, gx2, gv2, gC2, gF2, grid_v, adj, gx, gv, gC, gF, gstress, box=None,
                            hidden: int = 64) -> None:
    """adjoint gather of the substep + elasticity backward in one kernel (neuma_constitutive_bwd_gather)."""
    _ctx(x).call("neuma_constitutive_bwd_gather", kind, x.shape[0], hidden, ptr(staged), ptr(gWpart), C.byref(c),
                 ptr(x), ptr(v), ptr(Cm), ptr(F), ptr(stress), ptr(vol), ptr(rho), ptr(clip_bound), ptr(enabled),
                 ptr(v2), ptr(C2), ptr(gx2), ptr(gv2), ptr(gC2), ptr(gF2), ptr(box), ptr(grid_v), ptr(adj), ptr(gx),
                 ptr(gv), ptr(gC), ptr(gF), ptr(gstress), _stream(x))


def constitutive_svd_cache(like: torch.Tensor, v_out: Optional[torch.Tensor] = None,
                           v_in: Optional[torch.Tensor] = None) -> None:
    """The NEXT constitutive launch on `like`'s device writes the right singular vectors of its input
    rows to v_out [n, 3, 3] (forward kernels) / reads them from v_in and skips the Jacobi iteration
    (backward kernels) -- neuma_constitutive_svd_cache."""
    _ctx(like).call("neuma_constitutive_svd_cache", ptr(v_out), ptr(v_in))


def constitutive_reduce_partials(gWpart, gW0, gW1, gW2, hidden: int = 64) -> None:
    """gW* += sum of the slots; the slots are zeroed."""
    _ctx(gWpart).call("neuma_constitutive_reduce_partials", hidden, ptr(gWpart), ptr(gW0), ptr(gW1), ptr(gW2),
                      _stream(gWpart))


def mpm_planes_add(dst: torch.Tensor, src: torch.Tensor) -> None:
    """dst += src for packed float4 grid planes (slab-halo receive side)."""
    if dst.numel() != src.numel() or dst.numel() % 4:
        raise NeumaError("mpm_planes_add: shapes differ / not float4 nodes")
    _ctx(dst).call("neuma_mpm_planes_add", ptr(dst), ptr(src), C.c_size_t(dst.numel() // 4), _stream(dst))


def mpm_bwd_scatter(c: MPMConstC, x, F, clip_bound, enabled, v2, gx2, gv2, gC2, gF2, adj, box=None) -> None:
    _ctx(x).call("neuma_mpm_bwd_scatter", C.byref(c), x.shape[0], ptr(x), ptr(F), ptr(clip_bound), ptr(enabled),
                 ptr(v2), ptr(gx2), ptr(gv2), ptr(gC2), ptr(gF2), ptr(box), ptr(adj), _stream(x))


def mpm_bwd_node(c: MPMConstC, grid_v, adj, box) -> None:
    _ctx(adj).call("neuma_mpm_bwd_node", C.byref(c), ptr(grid_v), ptr(adj), ptr(box), _stream(adj))


def mpm_bwd_node_gather(c: MPMConstC, x, v, Cm, F, stress, vol, rho, clip_bound, enabled, v2, C2, gx2, gv2, gC2,
                        gF2, grid_v, adj, box, gx, gv, gC, gF, gstress) -> None:
    _ctx(x).call("neuma_mpm_bwd_node_gather", C.byref(c), x.shape[0], ptr(x), ptr(v), ptr(Cm), ptr(F),
                 ptr(stress), ptr(vol), ptr(rho), ptr(clip_bound), ptr(enabled), ptr(v2), ptr(C2), ptr(gx2),
                 ptr(gv2), ptr(gC2), ptr(gF2), ptr(grid_v), ptr(adj), ptr(box), ptr(gx), ptr(gv), ptr(gC), ptr(gF),
                 ptr(gstress), _stream(x))


def bind_build(means: torch.Tensor, cov6: torch.Tensor, x: torch.Tensor, threshold: float, max_particles: int = 10):
    """(counts[K], cols[K, max_particles] (-1 padded, ascending distance), pvals) -- see neuma_bind_build."""
    means, cov6, x = map(_f32, (means, cov6, x))
    K = means.shape[0]
    counts = torch.empty(K, dtype=torch.int32, device=x.device)
    cols = torch.empty(K, max_particles, dtype=torch.int32, device=x.device)
    pvals = torch.empty(K, max_particles, dtype=torch.float32, device=x.device)
    _ctx(x).call("neuma_bind_build", K, x.shape[0], ptr(means), ptr(cov6), ptr(x), C.c_float(threshold),
                 int(max_particles), ptr(counts), ptr(cols), ptr(pvals), _stream(x))
    return counts, cols, pvals


def _host3(t, default):
    vals = [float(v) for v in (t.detach().cpu().reshape(-1).tolist() if torch.is_tensor(t) else (t if t is not None else [default] * 3))]
    return (C.c_float * 3)(*vals)


def particles_to_gaussians(rowptr, col, w, slot, x, F, center, size, de_x_prev, k_prev, cov6, want_Fk: bool = False):
    """means3D [K,3], deformed cov [K,6] (, F_k [K,3,3]) from the particle state in one launch."""
    K = rowptr.shape[0] - 1
    means3D = torch.empty(K, 3, device=x.device, dtype=torch.float32)
    cov = torch.empty(K, 6, device=x.device, dtype=torch.float32)
    Fk = torch.empty(K, 3, 3, device=x.device, dtype=torch.float32) if want_Fk else None
    _ctx(x).call("neuma_particles_to_gaussians", K, ptr(rowptr), ptr(col), ptr(w), ptr(slot), ptr(x), ptr(F),
                 _host3(center, 0.0), _host3(size, 1.0), ptr(de_x_prev), ptr(k_prev), ptr(cov6), ptr(means3D), ptr(cov),
                 ptr(Fk), _stream(x))
    return means3D, cov, Fk


def gaussians_to_particles(rowptr, col, w, slot, g_means3D, size, gx_out) -> None:
    """gx_out [Np,3] += B^T g_means3D / size (rows through `slot`); the caller zero-fills gx_out."""
    K = rowptr.shape[0] - 1
    _ctx(g_means3D).call("neuma_gaussians_to_particles", K, ptr(rowptr), ptr(col), ptr(w), ptr(slot), ptr(g_means3D),
                         _host3(size, 1.0), ptr(gx_out), _stream(g_means3D))


def halo_push(src, peer_mailbox, peer_data_flag, my_ack_flag, my_seq, done) -> None:
    _ctx(src).call("neuma_halo_push", ptr(src), ptr(peer_mailbox), C.c_size_t(src.numel() // 4), ptr(peer_data_flag),
                   ptr(my_ack_flag), ptr(my_seq), ptr(done), _stream(src))


def halo_add_ack(dst, mailbox, my_data_flag, peer_ack_flag, my_seq, done) -> None:
    _ctx(dst).call("neuma_halo_add_ack", ptr(dst), ptr(mailbox), C.c_size_t(dst.numel() // 4), ptr(my_data_flag),
                   ptr(peer_ack_flag), ptr(my_seq), ptr(done), _stream(dst))


def halo_side(planes, peer_mail, peer_hdr, my_mail, my_hdr, peer_data_flag, my_data_flag, peer_ack_flag,
              my_ack_flag, npl: int, z0: int) -> HaloSideC:
    h = HaloSideC()
    for k, t in (("planes", planes), ("peer_mail", peer_mail), ("peer_hdr", peer_hdr), ("my_mail", my_mail),
                 ("my_hdr", my_hdr), ("peer_data_flag", peer_data_flag), ("my_data_flag", my_data_flag),
                 ("peer_ack_flag", peer_ack_flag), ("my_ack_flag", my_ack_flag)):
        setattr(h, k, t.data_ptr())
    h.npl, h.z0 = int(npl), int(z0)
    return h


def halo_exchange(c: MPMConstC, lo: Optional[HaloSideC], hi: Optional[HaloSideC], box, cnt) -> None:
    """ONE kernel: push my window of the shared planes to both neighbours, add theirs (neuma_halo_exchange)."""
    _ctx(box).call("neuma_halo_exchange", C.byref(c), C.byref(lo) if lo is not None else None,
                   C.byref(hi) if hi is not None else None, ptr(box), ptr(cnt), _stream(box))
