"""CPU checks of the MPM oracle: invariants, finite differences, edge cases."""
import itertools

import pytest
import torch

from oracle import mpm_oracle as O
from util import make_cloud


def test_weights_partition_of_unity():
    c = O.MPMConst(num_grids=16, dt=1e-3)
    x = torch.rand(1000, 3, dtype=torch.float64) * 0.8 + 0.1
    _, _, w = O._stencil(c, x)
    assert torch.allclose(w.sum(-1), torch.ones(1000, 3, dtype=torch.float64), atol=1e-12)


def test_base_is_truncation_not_floor():
    c = O.MPMConst(num_grids=16, dt=1e-3)
    x = torch.tensor([[0.01, 0.5, 0.5]], dtype=torch.float32)  # 0.16 - 0.5 = -0.34 -> int() = 0
    base, fx, _ = O._stencil(c, x)
    assert base[0, 0].item() == 0
    assert fx[0, 0].item() == pytest.approx(0.16, abs=1e-6)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_p2g_conserves_mass_and_momentum(dtype):
    c = O.MPMConst(num_grids=32, dt=1e-3)
    d = make_cloud(5000, 32, seed=1, dtype=dtype)
    zero = torch.zeros_like(d["C"])
    mv, m = O.p2g(c, d["x"], d["v"], zero, zero, d["vol"], d["rho"], d["enabled"])
    pm = d["vol"] * d["rho"]
    tol = 1e-4 if dtype == torch.float32 else 1e-10
    assert abs(m.sum().item() - pm.sum().item()) / pm.sum().item() < tol
    pmom = (pm[:, None] * d["v"]).sum(0)
    assert torch.allclose(mv.sum((0, 1, 2)), pmom, rtol=tol * 10, atol=tol)


def test_grid_op_empty_nodes_get_gravity_and_walls():
    for bc in ("noslip", "freeslip"):
        c = O.MPMConst(num_grids=8, dt=1e-2, bc=bc, gravity=(1.0, -9.8, 0.5))
        mv = torch.zeros(8, 8, 8, 3, dtype=torch.float64)
        m = torch.zeros(8, 8, 8, dtype=torch.float64)
        v = O.grid_op(c, mv, m)
        g = torch.tensor(c.gravity, dtype=torch.float64) * c.dt
        assert torch.allclose(v[3, 3, 3], g)
        # floor (py = 0), gravity points down: noslip zeroes all, freeslip only y
        if bc == "noslip":
            assert torch.all(v[3, 0, 3] == 0)
        else:
            assert v[3, 0, 3, 1] == 0 and v[3, 0, 3, 0] == g[0]
        # x-high wall with +x gravity
        assert v[7, 3, 3, 0] == 0


def test_disabled_particles_untouched():
    c = O.MPMConst(num_grids=16, dt=1e-3)
    d = make_cloud(300, 16, seed=3, dtype=torch.float64, disabled_frac=0.3)
    x2, v2, C2, F2 = O.substep_fwd(c, d["x"], d["v"], d["C"], d["F"], d["stress"], d["vol"], d["rho"],
                                   d["clip"], d["enabled"])
    off = d["enabled"] == 0
    assert off.any()
    assert torch.equal(x2[off], d["x"][off]) and torch.equal(F2[off], d["F"][off])
    assert torch.all(v2[off] == 0) and torch.all(C2[off] == 0)


def test_empty_input():
    c = O.MPMConst(num_grids=8, dt=1e-3)
    d = make_cloud(0, 8, dtype=torch.float64)
    out = O.substep_fwd(c, d["x"], d["v"], d["C"], d["F"], d["stress"], d["vol"], d["rho"], d["clip"],
                        d["enabled"])
    assert out[0].shape == (0, 3)


@pytest.mark.parametrize("bc", ["noslip", "freeslip"])
def test_autograd_adjoint_matches_finite_differences(bc):
    """The oracle adjoint (autograd) against fp64 central differences of the oracle forward."""
    torch.manual_seed(0)
    c = O.MPMConst(num_grids=8, dt=2e-3, bc=bc)
    d = make_cloud(40, 8, seed=5, lo=0.3, hi=0.7, dtype=torch.float64, stress_scale=5.0)
    names = ["x", "v", "C", "F", "stress"]
    ins = [d[k] for k in names]
    gouts = [torch.randn(40, 3, dtype=torch.float64), torch.randn(40, 3, dtype=torch.float64),
             torch.randn(40, 3, 3, dtype=torch.float64), torch.randn(40, 3, 3, dtype=torch.float64)]

    def loss(*a):
        outs = O.substep_fwd(c, *a, d["vol"], d["rho"], d["clip"], d["enabled"])
        return sum((o * g).sum() for o, g in zip(outs, gouts))

    grads = O.substep_bwd(c, *ins, d["vol"], d["rho"], d["clip"], d["enabled"], *gouts)
    h = 1e-6
    gen = torch.Generator().manual_seed(1)
    for k, (t, g) in enumerate(zip(ins, grads)):
        flat = t.reshape(-1)
        for idx in torch.randint(0, flat.numel(), (6,), generator=gen).tolist():
            tp, tm = t.clone().reshape(-1), t.clone().reshape(-1)
            tp[idx] += h
            tm[idx] -= h
            a_p = [tp.reshape(t.shape) if j == k else ins[j] for j in range(5)]
            a_m = [tm.reshape(t.shape) if j == k else ins[j] for j in range(5)]
            fd = (loss(*a_p) - loss(*a_m)) / (2 * h)
            an = g.reshape(-1)[idx]
            assert abs(fd - an) <= 1e-5 * max(1.0, abs(fd)), (names[k], idx, float(fd), float(an))
