"""GPU parity: the sm_100a rasterizer (through the C ABI) against the CPU oracle."""
import numpy as np
import pytest
import torch

from oracle import raster_oracle as RO
from scene import make_scene
from util import rel_err

pytestmark = pytest.mark.gpu


def _settings(s):
    from neuma_b200 import ops
    return ops.raster_settings(s.image_height, s.image_width, s.tanfovx, s.tanfovy, s.bg, 1.0,
                               s.viewmatrix, s.projmatrix, s.sh_degree, s.campos)


def _run_fwd(s, sc, dev, colors=None):
    from neuma_b200 import ops
    cs = _settings(s)
    img, radii = ops.raster_fwd(cs, sc["means"].to(dev), None if colors is not None else sc["shs"].to(dev),
                                None if colors is None else colors.to(dev), sc["opac"].to(dev),
                                sc["cov6"].to(dev))
    return cs, img, radii


@pytest.mark.parametrize("K,W,H,deg,seed", [(2000, 128, 128, 3, 0), (5000, 200, 152, 0, 1), (300, 70, 50, 2, 2)])
def test_binning_bit_exact(cuda_dev, K, W, H, deg, seed):
    """radii, sorted (tile << 32 | depth-bits) keys, Gaussian ids and tile ranges: exact."""
    from neuma_b200 import ops
    s, sc = make_scene(K, W=W, H=H, seed=seed, sh_degree=deg)
    cs, img, radii = _run_fwd(s, sc, cuda_dev)
    b = RO.bin_np(s, sc["means"].numpy(), sc["cov6"].numpy())
    assert np.array_equal(radii.cpu().numpy(), b["radii"])
    keys, ids, ranges = ops.raster_binning(img, b["tiles_x"] * b["tiles_y"])
    assert keys.shape[0] == b["keys"].shape[0]
    assert np.array_equal(keys.cpu().numpy().view(np.uint64), b["keys"])
    assert np.array_equal(ids.cpu().numpy(), b["ids"])
    assert np.array_equal(ranges.cpu().numpy(), b["ranges"])


@pytest.mark.parametrize("K,W,H,deg,bg", [(3000, 128, 128, 3, (0, 0, 0)), (1500, 160, 96, 1, (1, 1, 1)),
                                           (800, 64, 64, 0, (0.2, 0.5, 0.7))])
def test_forward_image(cuda_dev, K, W, H, deg, bg):
    s, sc = make_scene(K, W=W, H=H, seed=K, sh_degree=deg, bg=bg)
    cs, img, radii = _run_fwd(s, sc, cuda_dev)
    ref, _, aux = RO.rasterize(s, sc["means"], sc["shs"], None, sc["opac"], sc["cov6"], return_aux=True)
    ok = ~aux["borderline"]
    assert ok.float().mean() > 0.97     # pixels with no Gaussian sitting on a discrete threshold
    diff = (img.cpu().double() - ref).abs().max(0).values
    assert float(diff[ok].max()) < 1e-4 * max(1.0, float(ref.abs().max()))
    # the rest may flip one 1/255-alpha contribution at most
    assert float(diff.max()) < 2e-2
    assert rel_err(img, ref) < 2e-2


def test_precomputed_colors_mask_mode(cuda_dev):
    """force_mask_data path: colors_precomp = ones (tune/utils.py:390-404)."""
    s, sc = make_scene(1000, W=96, H=96, seed=9)
    cols = torch.ones(1000, 3)
    cs, img, radii = _run_fwd(s, sc, cuda_dev, colors=cols)
    ref, _, aux = RO.rasterize(s, sc["means"], None, cols, sc["opac"], sc["cov6"], return_aux=True)
    ok = ~aux["borderline"]
    diff = (img.cpu().double() - ref).abs().max(0).values
    assert float(diff[ok].max()) < 1e-4


@pytest.mark.parametrize("K,W,H,deg", [(1500, 96, 96, 3), (600, 64, 80, 0)])
def test_backward(cuda_dev, K, W, H, deg):
    from neuma_b200 import ops
    s, sc = make_scene(K, W=W, H=H, seed=K + 1, sh_degree=deg, bg=(0.1, 0.2, 0.3))
    cs, img, radii = _run_fwd(s, sc, cuda_dev)
    G = torch.randn(3, H, W, generator=torch.Generator().manual_seed(0))
    g3, g2, gcov, gop, gsh = ops.raster_bwd(cs, sc["means"].to(cuda_dev), sc["shs"].to(cuda_dev), None,
                                            sc["opac"].to(cuda_dev), sc["cov6"].to(cuda_dev), radii,
                                            G.to(cuda_dev), want_cov=True, want_opacity=True, want_color=True)
    leaves = [sc[k].double().requires_grad_(True) for k in ("means", "cov6", "opac", "shs")]
    binning = RO.bin_np(s, sc["means"].numpy(), sc["cov6"].numpy())
    ref = RO.render_torch(s, binning, leaves[0], leaves[3], None, leaves[2], leaves[1])
    og = torch.autograd.grad(ref, leaves, G.double())
    assert rel_err(g3, og[0]) < 2e-3
    assert rel_err(gcov, og[1]) < 2e-3
    assert rel_err(gop.reshape(-1), og[2].reshape(-1)) < 2e-3
    assert rel_err(gsh, og[3]) < 2e-3
    # relative L2 (robust to the rare threshold flip) at the 1e-4 level the north_star names
    for a, b in ((g3, og[0]), (gcov, og[1])):
        a, b = a.cpu().double(), b
        assert float((a - b).norm() / b.norm()) < 5e-4
    assert torch.all(g2[:, 2] == 0)


def test_empty_and_culled(cuda_dev):
    from neuma_b200 import ops
    s, sc = make_scene(50, W=64, H=64, seed=4, bg=(0.3, 0.3, 0.3))
    sc["means"][:, 2] = 5.0            # behind the camera (camera at z = 2.5 looking down -z)
    cs, img, radii = _run_fwd(s, sc, cuda_dev)
    assert torch.all(radii == 0)
    assert torch.allclose(img, torch.full_like(img, 0.3))
    g3, g2, *_ = ops.raster_bwd(cs, sc["means"].to(cuda_dev), sc["shs"].to(cuda_dev), None,
                                sc["opac"].to(cuda_dev), sc["cov6"].to(cuda_dev), radii, torch.ones_like(img))
    assert torch.all(g3 == 0)
    empty = {k: v[:0] for k, v in sc.items()}
    cs, img, radii = _run_fwd(s, empty, cuda_dev)
    assert torch.allclose(img, torch.full_like(img, 0.3)) and radii.numel() == 0


def test_full_size_800(cuda_dev):
    """BASELINE config-2 shape: 25k Gaussians, 800x800, sh 3 -- size-independent properties."""
    from neuma_b200 import ops
    s, sc = make_scene(25000, W=800, H=800, seed=42, dx=1 / 64)
    cs, img, radii = _run_fwd(s, sc, cuda_dev)
    assert torch.isfinite(img).all() and float(img.min()) >= 0.0
    keys, ids, ranges = ops.raster_binning(img, 50 * 50)
    k = keys.cpu().numpy().view(np.uint64)
    assert np.all(k[1:] >= k[:-1])                                    # sortedness
    r = ranges.cpu().numpy()
    assert int((r[:, 1] - r[:, 0]).sum()) == k.shape[0]               # ranges partition the list
    # linearity of the backward in dL/dimage
    G = torch.randn_like(img)
    a = ops.raster_bwd(cs, sc["means"].to(cuda_dev), sc["shs"].to(cuda_dev), None, sc["opac"].to(cuda_dev),
                       sc["cov6"].to(cuda_dev), radii, G)[0]
    b = ops.raster_bwd(cs, sc["means"].to(cuda_dev), sc["shs"].to(cuda_dev), None, sc["opac"].to(cuda_dev),
                       sc["cov6"].to(cuda_dev), radii, 2.0 * G)[0]
    assert float((b - 2 * a).norm() / b.norm()) < 1e-4


def test_saved_state_backward(cuda_dev):
    """Two forwards, then the backward of the FIRST from its saved state (rollout order,
    finetune.py:360-414) == its backward taken immediately."""
    from neuma_b200 import ops
    s, a = make_scene(1200, W=96, H=96, seed=21)
    _, b = make_scene(1500, W=96, H=96, seed=22)
    G = torch.randn(3, 96, 96).to(cuda_dev)
    cs, img, radii = _run_fwd(s, a, cuda_dev)
    args = (a["means"].to(cuda_dev), a["shs"].to(cuda_dev), None, a["opac"].to(cuda_dev), a["cov6"].to(cuda_dev))
    ref = ops.raster_bwd(cs, *args, radii, G)[0]
    saved = ops.raster_save_state(img)
    _run_fwd(s, b, cuda_dev)                       # clobbers the ctx state
    got = ops.raster_bwd(cs, *args, radii, G, saved=saved)[0]
    assert rel_err(got, ref) < 1e-5
