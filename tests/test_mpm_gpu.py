"""GPU parity: the sm_100a MPM kernels (through the C ABI) against the CPU oracle."""
import pytest
import torch

from oracle import mpm_oracle as O
from util import make_cloud, rel_err

pytestmark = pytest.mark.gpu

RTOL = 1e-4  # north_star: particle x/v/F within 1e-4 rel fp32


def _dev(d, dev):
    return {k: t.to(dev) for k, t in d.items()}


def _consts(G, dt, bc, eps=6e-7, gravity=(0.0, -9.8, 0.0), bound=1):
    from neuma_b200 import ops
    return (O.MPMConst(num_grids=G, dt=dt, bound=bound, gravity=gravity, eps=eps, bc=bc),
            ops.mpm_const(G, dt, bound, gravity, eps, bc))


@pytest.mark.parametrize("n,G,bc", [(5000, 32, "noslip"), (20000, 64, "freeslip"), (333, 16, "noslip")])
def test_p2g_gridop_g2p_stagewise(cuda_dev, n, G, bc):
    from neuma_b200 import ops
    co, cc = _consts(G, 1e-3, bc)
    d = make_cloud(n, G, seed=n, disabled_frac=0.1)
    g = _dev(d, cuda_dev)
    ops.mpm_grid_clear(cc, g["x"])
    ops.mpm_p2g(cc, g["x"], g["v"], g["C"], g["stress"], g["vol"], g["rho"], g["enabled"])
    ops.mpm_grid_op(cc, g["x"])
    mv, m, v = ops.mpm_grid_export(cc, g["x"], want_v=True)
    d64 = {k: (t.double() if t.is_floating_point() else t) for k, t in d.items()}
    mv_o, m_o = O.p2g(co, d64["x"], d64["v"], d64["C"], d64["stress"], d64["vol"], d64["rho"], d["enabled"])
    assert rel_err(m, m_o) < RTOL
    assert rel_err(mv, mv_o) < RTOL
    # conservation on the device result
    pm = (d64["vol"] * d64["rho"])[d["enabled"] != 0]
    assert abs(m.double().sum().item() - pm.sum().item()) / pm.sum().item() < 1e-4
    # grid_op on the DEVICE grid values so that the m > 0 branch is identical
    v_o = O.grid_op(co, mv.cpu().double(), m.cpu().double())
    assert rel_err(v, v_o) < RTOL
    # g2p, both the dense-grid and the lazy-grid_op variants
    for lazy in (False, True):
        x2, v2, C2, F2 = ops.mpm_g2p(cc, g["x"], g["F"], g["clip"], g["enabled"], lazy_grid_op=lazy)
        torch.cuda.synchronize()
        ox2, ov2, oC2, oF2 = O.g2p(co, d64["x"], d64["F"], d64["clip"], d["enabled"], v_o)
        on = (d["enabled"] != 0)
        assert rel_err(x2.cpu()[on], ox2[on]) < RTOL
        assert rel_err(v2.cpu()[on], ov2[on]) < RTOL
        assert rel_err(C2.cpu()[on], oC2[on]) < RTOL
        assert rel_err(F2.cpu()[on], oF2[on]) < RTOL


@pytest.mark.parametrize("n,G,bc,dt", [(5000, 32, "noslip", 1e-3), (50000, 64, "noslip", 1e-3),
                                       (3000, 16, "freeslip", 5e-4)])
def test_substep_fwd_bwd(cuda_dev, n, G, bc, dt):
    from neuma_b200 import ops
    co, cc = _consts(G, dt, bc)
    d = make_cloud(n, G, seed=7 + n, disabled_frac=0.05)
    g = _dev(d, cuda_dev)
    x2, v2, C2, F2 = ops.mpm_substep_fwd(cc, g["x"], g["v"], g["C"], g["F"], g["stress"], g["vol"],
                                         g["rho"], g["clip"], g["enabled"])
    d64 = {k: (t.double() if t.is_floating_point() else t) for k, t in d.items()}
    o = O.substep_fwd(co, d64["x"], d64["v"], d64["C"], d64["F"], d64["stress"], d64["vol"],
                      d64["rho"], d64["clip"], d["enabled"])
    on = d["enabled"] != 0
    for a, b, name in zip((x2, v2, C2, F2), o, "xvCF"):
        assert rel_err(a.cpu()[on], b[on]) < RTOL, name
    gen = torch.Generator().manual_seed(11)
    gouts = [torch.randn(n, 3, generator=gen), torch.randn(n, 3, generator=gen),
             torch.randn(n, 3, 3, generator=gen), torch.randn(n, 3, 3, generator=gen)]
    grads = ops.mpm_substep_bwd(cc, g["x"], g["v"], g["C"], g["F"], g["stress"], g["vol"], g["rho"],
                                g["clip"], g["enabled"], v2, *[t.to(cuda_dev) for t in gouts])
    og = O.substep_bwd(co, d["x"], d["v"], d["C"], d["F"], d["stress"], d64["vol"], d64["rho"],
                       d64["clip"], d["enabled"], *gouts)
    for a, b, name in zip(grads, og, ["gx", "gv", "gC", "gF", "gstress"]):
        assert rel_err(a, b) < 5e-4, name  # fp32 atomics vs fp64 autograd
        assert torch.all(a[~on.to(cuda_dev)] == 0)


def test_in_place_forward_sim(cuda_dev):
    """MPMForwardSim aliases state_curr and state_next (interface.py:133)."""
    from neuma_b200 import ops
    co, cc = _consts(32, 1e-3, "noslip")
    d = make_cloud(4000, 32, seed=3)
    g = _dev(d, cuda_dev)
    ref = ops.mpm_substep_fwd(cc, g["x"], g["v"], g["C"], g["F"], g["stress"], g["vol"], g["rho"],
                              g["clip"], g["enabled"])
    x, v, C, F = g["x"].clone(), g["v"].clone(), g["C"].clone(), g["F"].clone()
    ops.mpm_substep_fwd(cc, x, v, C, F, g["stress"], g["vol"], g["rho"], g["clip"], g["enabled"],
                        out=(x, v, C, F))
    for a, b in zip((x, v, C, F), ref):
        assert rel_err(a, b) < 1e-5


def test_empty_and_clamped(cuda_dev):
    from neuma_b200 import ops
    co, cc = _consts(16, 1e-3, "noslip")
    d = make_cloud(0, 16)
    g = _dev(d, cuda_dev)
    out = ops.mpm_substep_fwd(cc, g["x"], g["v"], g["C"], g["F"], g["stress"], g["vol"], g["rho"],
                              g["clip"], g["enabled"])
    assert out[0].shape == (0, 3)
    # particles pushed against the clamp: x' must equal the clip bound exactly
    d = make_cloud(64, 16, seed=2, lo=0.02, hi=0.05)
    d["v"][:] = torch.tensor([-40.0, -40.0, -40.0])
    g = _dev(d, cuda_dev)
    x2, *_ = ops.mpm_substep_fwd(cc, g["x"], g["v"], g["C"] * 0, g["F"], g["stress"] * 0, g["vol"],
                                 g["rho"], g["clip"], g["enabled"])
    assert torch.all(x2 >= 0.1 / 16 - 1e-9)


def test_twenty_substep_rollout(cuda_dev):
    """BASELINE config 1 shape: 5k particles / 32^3 / 20 substeps, forward only."""
    from neuma_b200 import ops
    co, cc = _consts(32, 1e-3, "noslip")
    d = make_cloud(5000, 32, seed=42, stress_scale=0.0)
    d["C"] *= 0
    d["F"] = torch.eye(3).expand(5000, 3, 3).contiguous()
    g = _dev(d, cuda_dev)
    x, v, C, F = g["x"].clone(), g["v"].clone(), g["C"].clone(), g["F"].clone()
    ox, ov, oC, oF = d["x"].double(), d["v"].double(), d["C"].double(), d["F"].double()
    for _ in range(20):
        ops.mpm_substep_fwd(cc, x, v, C, F, g["stress"], g["vol"], g["rho"], g["clip"], g["enabled"],
                            out=(x, v, C, F))
        ox, ov, oC, oF = O.substep_fwd(co, ox, ov, oC, oF, d["stress"].double(), d["vol"].double(),
                                       d["rho"].double(), d["clip"].double(), d["enabled"])
    assert rel_err(x, ox) < RTOL and rel_err(v, ov) < RTOL and rel_err(F, oF) < RTOL
