"""The MPM / SVD / deform_cov oracle and the CUDA kernels against golden vectors made by executing
the reference's own kernel source under a Warp-semantics shim (tests/golden/make_golden_mpm.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import constitutive_oracle as CO
from oracle import mpm_oracle as O
from oracle import raster_oracle as RO
from util import rel_err

Z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mpm_shim.npz"))


def case(tag):
    g = {k[2:]: Z[k] for k in Z.files if k.startswith(tag + "_")}
    c = O.MPMConst(num_grids=int(g["G"]), dt=float(g["dt"]), bound=int(g["bound"]),
                   gravity=tuple(float(v) for v in g["gravity"]), eps=float(g["eps"]), bc=str(g["bc"]))
    t = {k: torch.from_numpy(np.asarray(g[k])) for k in ("x", "v", "C", "F", "stress", "vol", "rho", "clip", "enabled",
                                                         "grid_mv", "grid_m", "grid_v", "x2", "v2", "C2", "F2")}
    return c, t


@pytest.mark.parametrize("tag", ["a", "b"])
def test_oracle_matches_reference_kernel_source(tag):
    c, t = case(tag)
    mv, m = O.p2g(c, t["x"], t["v"], t["C"], t["stress"], t["vol"], t["rho"], t["enabled"])
    # per-element metric (tests/util.py): fp32 sums in a different order, a few ulp per node
    assert rel_err(m, t["grid_m"]) < 2e-6 and rel_err(mv, t["grid_mv"]) < 1e-5
    # grid_op on the reference's own (mv, m) so that the m > 0 / wall branches see identical inputs
    v = O.grid_op(c, t["grid_mv"], t["grid_m"])
    assert rel_err(v, t["grid_v"]) < 1e-6
    x2, v2, C2, F2 = O.g2p(c, t["x"], t["F"], t["clip"], t["enabled"], t["grid_v"],
                           x_prev=torch.zeros_like(t["x"]), F_prev=torch.eye(3).expand_as(t["F"]).contiguous())
    for a, b in ((x2, t["x2"]), (v2, t["v2"]), (C2, t["C2"]), (F2, t["F2"])):
        assert rel_err(a, b) < 2e-6
    # whole substep
    out = O.substep_fwd(c, t["x"], t["v"], t["C"], t["F"], t["stress"], t["vol"], t["rho"], t["clip"], t["enabled"])
    on = t["enabled"] != 0
    for a, b in zip(out, (t["x2"], t["v2"], t["C2"], t["F2"])):
        assert rel_err(a[on], b[on]) < 2e-5


def test_svd_oracle_matches_batch_svd_source():
    A = torch.from_numpy(Z["svd_A"]).double()
    U, s, Vh = CO.svd_signfix(A)
    gU, gs, gVh = [torch.from_numpy(Z[k]).double() for k in ("svd_U", "svd_sigma", "svd_Vh")]
    assert torch.allclose(s, gs, atol=1e-5)
    assert torch.allclose(U @ Vh, gU @ gVh, atol=1e-4)            # rotation-invariant comparison
    assert torch.allclose(gU @ torch.diag_embed(gs) @ gVh, A, atol=1e-5)
    assert torch.all(torch.linalg.det(gU) > 0) and torch.all(torch.linalg.det(gVh) > 0)


def test_deform_cov_oracle_matches_kernel_source():
    out = RO.deform_cov(torch.from_numpy(Z["cov_in"]).double(), torch.from_numpy(Z["cov_F"]).double())
    assert rel_err(out, torch.from_numpy(Z["cov_out"])) < 2e-6


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["a", "b"])
def test_cuda_matches_reference_kernel_source(cuda_dev, tag):
    from neuma_b200 import ops
    c, t = case(tag)
    cc = ops.mpm_const(c.num_grids, c.dt, c.bound, c.gravity, c.eps, c.bc)
    g = {k: v.to(cuda_dev) for k, v in t.items()}
    ops.mpm_grid_clear(cc, g["x"])
    ops.mpm_p2g(cc, g["x"], g["v"], g["C"], g["stress"], g["vol"], g["rho"], g["enabled"])
    ops.mpm_grid_op(cc, g["x"])
    mv, m, v = ops.mpm_grid_export(cc, g["x"], want_v=True)
    assert rel_err(m, t["grid_m"]) < 1e-5 and rel_err(mv, t["grid_mv"]) < 1e-5 and rel_err(v, t["grid_v"]) < 1e-5
    x2, v2, C2, F2 = ops.mpm_substep_fwd(cc, g["x"], g["v"], g["C"], g["F"], g["stress"], g["vol"], g["rho"],
                                         g["clip"], g["enabled"])
    on = (t["enabled"] != 0)
    for a, b in zip((x2, v2, C2, F2), (t["x2"], t["v2"], t["C2"], t["F2"])):
        assert rel_err(a.cpu()[on], b[on]) < 1e-4


@pytest.mark.gpu
def test_cuda_svd_and_cov_match_reference_kernel_source(cuda_dev):
    from neuma_b200 import ops
    A = torch.from_numpy(Z["svd_A"]).to(cuda_dev)
    U, s, Vh = [x.cpu().double() for x in ops.svd3_fwd(A)]
    gU, gs, gVh = [torch.from_numpy(Z[k]).double() for k in ("svd_U", "svd_sigma", "svd_Vh")]
    assert torch.allclose(s, gs, atol=2e-5)
    assert torch.allclose(U @ Vh, gU @ gVh, atol=2e-4)
    out = ops.deform_cov(torch.from_numpy(Z["cov_in"]).to(cuda_dev), torch.from_numpy(Z["cov_F"]).to(cuda_dev))
    assert rel_err(out, torch.from_numpy(Z["cov_out"])) < 1e-5
