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
