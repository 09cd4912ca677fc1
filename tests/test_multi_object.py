"""Multi-object path: ComposeMaterial section dispatch (preset.py:12-27), per-step
`enabled` spans (mpm.py:67-72) and preprocess_for_rasterization (tune/utils.py:475-523)."""
from types import SimpleNamespace

import pytest
import torch

from util import make_cloud, rel_err


def test_compose_material_dispatch_cpu():
    """host logic only: slices, empty sections, order -- with stub materials"""
    from neuma_b200.material import ComposeMaterial

    class Scale(torch.nn.Module):
        def __init__(self, k):
            super().__init__()
            self.k = k

        def forward(self, F):
            return self.k * F

    F = torch.arange(10 * 9, dtype=torch.float32).reshape(10, 3, 3)
    cm = ComposeMaterial([Scale(1.0), Scale(2.0), Scale(3.0)], [4, 0, 6])
    out = cm(F)
    assert torch.equal(out[:4], F[:4]) and torch.equal(out[4:], 3.0 * F[4:])
    cm.update_sections([2, 3, 5])
    out = cm(F)
    assert torch.equal(out[2:5], 2.0 * F[2:5]) and torch.equal(out[5:], 3.0 * F[5:])
    with pytest.raises(RuntimeError):
        ComposeMaterial([Scale(1.0)], [3])(F)          # sections must cover all particles (torch.split)


@pytest.mark.gpu
def test_compose_material_matches_per_object_nets(cuda_dev):
    import numpy as np, os
    from neuma_b200 import ops
    from neuma_b200.material import ComposeMaterial
    from oracle import constitutive_oracle as CO
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

    class Net(torch.nn.Module):
        def __init__(self, W):
            super().__init__()
            self.W = [w.float().to(cuda_dev) for w in W]

        def forward(self, F):
            return ops.constitutive_fwd(ops.KIND_ELASTIC, F, *self.W)

    Ws = []
    for name in ("jelly", "sand"):
        z = np.load(os.path.join(root, "tests", "golden", f"material_{name}.npz"))
        Ws.append([torch.from_numpy(z[f"param_elastic.{l}.fc.weight"]).double() for l in ("layers.0", "layers.1", "final_layer")])
    torch.manual_seed(2)
    F = torch.eye(3, dtype=torch.float64) + 0.05 * torch.randn(300, 3, 3, dtype=torch.float64)
    cm = ComposeMaterial([Net(Ws[0]), Net(Ws[1])], [130, 170])
    out = cm(F.float().to(cuda_dev))
    ref = torch.cat([CO.elasticity(F[:130], Ws[0]), CO.elasticity(F[130:], Ws[1])])
    assert rel_err(out, ref) < 1e-4


@pytest.mark.gpu
def test_enabled_spans_and_preprocess_for_rasterization(cuda_dev):
    from neuma_b200 import ops
    from neuma_b200.glue import Bindings, preprocess_for_rasterization
    from neuma_b200.sim.mpm import MPMStatics
    # ---- update_enabled: object 0 lives in steps [0, 5), object 1 in [3, 10)
    st = MPMStatics()
    st.init(10, cuda_dev)
    st.update_enabled([4, 6], [(0, 5), (3, 10)], step=4)
    assert st.enabled.tolist() == [1] * 10
    st.update_enabled([4, 6], [(0, 5), (3, 10)], step=7)
    assert st.enabled.tolist() == [0] * 4 + [1] * 6
    # disabled particles neither scatter nor move (the reference's kernels return early)
    d = {k: t.to(cuda_dev) for k, t in make_cloud(10, 16, seed=1).items()}
    cc = ops.mpm_const(16, 1e-3, 1, (0.0, -9.8, 0.0), 6e-7, "noslip")
    x2, *_ = ops.mpm_substep_fwd(cc, d["x"], d["v"], d["C"], d["F"], d["stress"], d["vol"], d["rho"], d["clip"],
                                 st.enabled, out=tuple(t.clone() for t in (d["x"], d["v"], d["C"], d["F"])))
    assert torch.equal(x2[:4], d["x"][:4]) and not torch.equal(x2[4:], d["x"][4:])
    # ---- preprocess_for_rasterization on two objects == torch.sparse.mm per object, concatenated
    gen = torch.Generator().manual_seed(0)
    objs = []
    for K, Np in ((20, 50), (30, 70)):
        rows = torch.arange(K).repeat_interleave(3)
        cols = torch.randint(0, Np, (K * 3,), generator=gen)
        B = torch.sparse_coo_tensor(torch.stack([rows, cols]), torch.full((K * 3,), 1 / 3.0), (K, Np)).coalesce()
        objs.append(dict(B=B, p_prev=torch.rand(Np, 3, generator=gen), p_curr=torch.rand(Np, 3, generator=gen),
                         k_prev=torch.rand(K, 3, generator=gen), F=torch.randn(Np, 3, 3, generator=gen),
                         g=SimpleNamespace(get_covariance=lambda scaling_modifier=1.0, K=K: torch.full((K, 6), float(scaling_modifier), device=cuda_dev),
                                           get_opacity=torch.ones(K, 1, device=cuda_dev), get_features=torch.zeros(K, 16, 3, device=cuda_dev),
                                           active_sh_degree=3)))
    dev = lambda t: t.to(cuda_dev)
    pack = preprocess_for_rasterization([o["g"] for o in objs], [dev(o["F"]) for o in objs], [dev(o["k_prev"]) for o in objs],
                                        [dev(o["p_curr"]) for o in objs], [dev(o["p_prev"]) for o in objs],
                                        [Bindings.from_sparse(dev(o["B"])) for o in objs], [1.0, 2.0])
    ref_x = torch.cat([o["k_prev"] + torch.sparse.mm(o["B"], o["p_curr"] - o["p_prev"]) for o in objs])
    ref_F = torch.cat([torch.sparse.mm(o["B"], o["F"].reshape(-1, 9)).reshape(-1, 3, 3) for o in objs])
    assert rel_err(pack["means3D"], ref_x) < 1e-5 and rel_err(pack["deform_grad"], ref_F) < 1e-5
    assert pack["cov3D"].shape == (50, 6) and float(pack["cov3D"][25, 0]) == 2.0 and pack["active_sh_degree"] == 3
