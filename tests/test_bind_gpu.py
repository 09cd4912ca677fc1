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
