"""GPU parity: binding SpMM (fwd + transpose) and covariance deformation."""
import pytest
import torch

from oracle import raster_oracle as RO
from scene import make_scene
from util import rel_err

pytestmark = pytest.mark.gpu


def make_bindings(K, Np, seed=0, max_nnz=10):
    g = torch.Generator().manual_seed(seed)
    counts = torch.randint(0, max_nnz + 1, (K,), generator=g)   # rows with 0 entries occur (pruned later upstream)
    rowptr = torch.zeros(K + 1, dtype=torch.int32)
    rowptr[1:] = torch.cumsum(counts, 0)
    nnz = int(rowptr[-1])
    col = torch.randint(0, Np, (nnz,), generator=g, dtype=torch.int32)
    w = torch.cat([torch.full((int(c),), 1.0 / int(c)) for c in counts if c > 0]) if nnz else torch.zeros(0)
    rows = torch.repeat_interleave(torch.arange(K), counts)
    B = torch.sparse_coo_tensor(torch.stack([rows, col.long()]), w.double(), (K, Np)).coalesce()
    return rowptr, col, w.float(), B


@pytest.mark.parametrize("K,Np", [(500, 1000), (25000, 50000), (1, 7)])
def test_bind_spmm(cuda_dev, K, Np):
    from neuma_b200 import ops
    rowptr, col, w, B = make_bindings(K, Np, seed=K)
    x = torch.randn(Np, 3)
    F = torch.randn(Np, 3, 3)
    d = cuda_dev
    kx = ops.bind_spmm(rowptr.to(d), col.to(d), w.to(d), x.to(d), Np)
    kF = ops.bind_spmm(rowptr.to(d), col.to(d), w.to(d), F.reshape(Np, 9).to(d), Np)
    assert rel_err(kx, torch.sparse.mm(B, x.double())) < 1e-5
    assert rel_err(kF.reshape(K, 3, 3), RO.bindings_F(F.double(), B)) < 1e-5
    g = torch.randn(K, 9)
    gt = ops.bind_spmm(rowptr.to(d), col.to(d), w.to(d), g.to(d), Np, transpose=True)
    assert rel_err(gt, torch.sparse.mm(B.t(), g.double())) < 1e-5


def test_deform_cov(cuda_dev):
    from neuma_b200 import ops
    s, sc = make_scene(5000, seed=1)
    F = torch.eye(3) + 0.2 * torch.randn(5000, 3, 3)
    out = ops.deform_cov(sc["cov6"].to(cuda_dev), F.to(cuda_dev))
    assert rel_err(out, RO.deform_cov(sc["cov6"].double(), F.double())) < 1e-5


def test_binding_construction_matches_reference_rule(cuda_dev):
    """gaussian_binding_with_clip_v1 (binding_utils.py:199-285): inside the chi2 ellipsoid, the <= 10
    nearest in Mahalanobis distance, uniform weights -- against a float64 restatement."""
    from types import SimpleNamespace
    from scipy.stats import chi2
    from neuma_b200 import synthetic as S
    from neuma_b200.glue import gaussian_binding_with_clip_v1
    torch.manual_seed(0)
    Np, K, G = 6000, 700, 32
    x = 0.3 + 0.4 * torch.rand(Np, 3)
    g = S.make_gaussians(x, K, G, seed=5)
    K = g["K"]
    gm = SimpleNamespace(get_xyz=g["xyz"].to(cuda_dev), get_covariance=lambda: g["cov6"].to(cuda_dev))
    B = gaussian_binding_with_clip_v1(gm, x.to(cuda_dev), confidence=0.95, max_particles=10)
    assert B.shape == (K, Np)
    rows, cols = B.indices().cpu()
    vals = B.values().cpu()
    thr = float(chi2.ppf(0.95, 3))
    c = g["cov6"].double()
    Sig = torch.stack([c[:, 0], c[:, 1], c[:, 2], c[:, 1], c[:, 3], c[:, 4], c[:, 2], c[:, 4], c[:, 5]], -1).reshape(K, 3, 3)
    Sinv = torch.linalg.inv(Sig)
    d = x.double()[None] - g["xyz"].double()[:, None]                   # [K, Np, 3]
    p = torch.einsum("kni,kij,knj->kn", d, Sinv, d)
    eps = 1e-4
    for k in range(K):
        sel = cols[rows == k]
        n_in = int((p[k] <= thr).sum())
        want = min(n_in, 10)
        # particles within eps of the threshold / of the 10th distance may fall either side in fp32
        slack = int(((p[k] - thr).abs() <= eps * thr).sum())
        assert abs(len(sel) - want) <= slack, (k, len(sel), want)
        assert torch.all(p[k][sel] <= thr * (1 + eps))
        if n_in > 10:
            kth = torch.sort(p[k]).values[9]
            assert torch.all(p[k][sel] <= kth * (1 + eps))
        assert torch.allclose(vals[rows == k], torch.full((len(sel),), 1.0 / len(sel)))


def test_fused_particles_to_gaussians_and_transpose(cuda_dev):
    """one launch == denormalise + compute_bindings_xyz + compute_bindings_F + deform_cov_by_F
    (tune/utils.py:424-472, simulation_utils.py:25-48), with the engine's cell-order indirection."""
    from neuma_b200 import ops, synthetic as S
    torch.manual_seed(1)
    Np, G = 5000, 32
    x = 0.3 + 0.4 * torch.rand(Np, 3)
    g = S.make_gaussians(x, 900, G, seed=7)
    K = g["K"]
    F = torch.eye(3) + 0.1 * torch.randn(Np, 3, 3)
    x_prev = x + 0.01 * torch.randn(Np, 3)
    center, size = torch.tensor([0.1, -0.2, 0.05]), torch.tensor([1.5, 0.7, 2.0])
    de_prev = (x_prev - center) / size
    perm = torch.randperm(Np)                       # engine row i holds caller particle perm[i]
    slot = torch.empty(Np, dtype=torch.int64)
    slot[perm] = torch.arange(Np)
    dev = lambda t: t.to(cuda_dev)
    means3D, cov, Fk = ops.particles_to_gaussians(dev(g["rowptr"]), dev(g["col"]), dev(g["w"]), dev(slot.int()),
                                                  dev(x[perm].contiguous()), dev(F[perm].contiguous()), center, size,
                                                  dev(de_prev), dev(g["xyz"]), dev(g["cov6"]), want_Fk=True)
    B = torch.sparse_csr_tensor(g["rowptr"].long(), g["col"].long(), g["w"].double(), (K, Np)).to_sparse_coo()
    de = ((x - center) / size).double()
    ref_m = g["xyz"].double() + torch.sparse.mm(B, de - de_prev.double())
    ref_F = torch.sparse.mm(B, F.double().reshape(Np, 9)).reshape(K, 3, 3)
    c = g["cov6"].double()
    Sig = torch.stack([c[:, 0], c[:, 1], c[:, 2], c[:, 1], c[:, 3], c[:, 4], c[:, 2], c[:, 4], c[:, 5]], -1).reshape(K, 3, 3)
    R = ref_F @ Sig @ ref_F.transpose(1, 2)
    ref_cov = torch.stack([R[:, 0, 0], R[:, 0, 1], R[:, 0, 2], R[:, 1, 1], R[:, 1, 2], R[:, 2, 2]], -1)
    assert rel_err(means3D, ref_m) < 1e-5 and rel_err(Fk, ref_F) < 1e-5 and rel_err(cov, ref_cov) < 1e-5
    # transpose: dL/dmeans3D -> dL/dx in engine rows
    gm = torch.randn(K, 3)
    gx = torch.zeros(Np, 3, device=cuda_dev)
    ops.gaussians_to_particles(dev(g["rowptr"]), dev(g["col"]), dev(g["w"]), dev(slot.int()), dev(gm), size, gx)
    ref_gx = torch.sparse.mm(B.t(), gm.double()) / size.double()
    assert rel_err(gx.cpu()[slot], ref_gx) < 1e-5
