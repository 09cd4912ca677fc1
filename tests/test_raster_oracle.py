"""CPU checks of the rasterizer oracle (what can be pinned without the external CUDA extension)."""
import os

import numpy as np
import pytest
import torch

from oracle import raster_oracle as RO
from scene import make_scene

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "d3gs_helpers.npz")


def test_sh_matches_reference_eval_sh():
    z = np.load(GOLD)
    dirs, sh = torch.from_numpy(z["sh_dirs"]), torch.from_numpy(z["sh_coeffs"])
    for deg in range(4):
        rgb = torch.clamp_min(RO.sh_to_rgb(deg, sh, dirs), 0.0)
        assert torch.allclose(rgb, torch.from_numpy(z[f"sh_rgb_deg{deg}"]), atol=1e-12)


def test_camera_matches_reference_graphics_utils():
    z = np.load(GOLD)
    s = RO.look_at_camera(W=800, H=800, fov_deg=40.0, eye=(0.5, 0.5, 2.5))
    assert np.allclose(s.viewmatrix, z["cam_world_view"], atol=1e-6)
    assert np.allclose(s.projmatrix, z["cam_full_proj"], atol=1e-5)
    assert np.allclose(s.campos, z["cam_center"], atol=1e-6)


def test_single_gaussian_center_pixel():
    s = RO.look_at_camera(W=64, H=64)
    s.bg = np.array([0.2, 0.4, 0.6], np.float32)
    # the pixel whose centre the Gaussian projects onto exactly: ndc 0 -> pix 31.5; use an offset mean
    means = torch.tensor([[0.5, 0.5, 0.5]], dtype=torch.float64)
    cov6 = torch.tensor([[1e-3, 0, 0, 1e-3, 0, 1e-3]], dtype=torch.float64)
    col = torch.tensor([[0.9, 0.5, 0.1]], dtype=torch.float64)
    op = torch.tensor([[0.8]], dtype=torch.float64)
    img, radii, aux = RO.rasterize(s, means, None, col, op, cov6, return_aux=True)
    assert radii[0] > 0
    pre = aux["pre"]
    px, py = float(pre["px"][0]), float(pre["py"][0])
    assert abs(px - 31.5) < 1e-6 and abs(py - 31.5) < 1e-6
    # nearest pixel (31, 31): d = (0.5, 0.5)
    A, B, C = float(pre["A"][0]), float(pre["B"][0]), float(pre["C"][0])
    power = -0.5 * (A * 0.25 + C * 0.25) - B * 0.25
    alpha = min(0.99, 0.8 * np.exp(power))
    expect = np.array([0.9, 0.5, 0.1]) * alpha + (1 - alpha) * s.bg
    assert np.allclose(img[:, 31, 31].numpy(), expect, atol=1e-9)
    # far corner: untouched background
    assert np.allclose(img[:, 0, 0].numpy(), s.bg, atol=1e-12)


def test_binning_is_sorted_and_ranges_partition():
    s, sc = make_scene(400, W=96, H=80, seed=3)
    b = RO.bin_np(s, sc["means"].numpy(), sc["cov6"].numpy())
    keys = b["keys"]
    assert np.all(keys[1:] >= keys[:-1])
    assert b["ranges"][:, 1].max() == keys.shape[0]
    covered = sum(int(e - st) for st, e in b["ranges"])
    assert covered == keys.shape[0]
    for t, (st, e) in enumerate(b["ranges"]):
        assert np.all((keys[st:e] >> np.uint64(32)) == t)


def test_early_termination_and_contributor_count():
    s = RO.look_at_camera(W=32, H=32)
    K = 40  # a stack of opaque Gaussians at the same pixel, front to back
    means = torch.stack([torch.full((K,), 0.5), torch.full((K,), 0.5), torch.linspace(1.0, 0.2, K)], -1).double()
    cov6 = torch.tensor([[4e-3, 0, 0, 4e-3, 0, 4e-3]], dtype=torch.float64).repeat(K, 1)
    col = torch.rand(K, 3, dtype=torch.float64)
    op = torch.full((K, 1), 0.9, dtype=torch.float64)
    img, radii, aux = RO.rasterize(s, means, None, col, op, cov6, return_aux=True)
    n = int(aux["n_contrib"][16, 16])
    assert 0 < n < K                      # stopped before exhausting the list
    assert float(aux["final_T"][16, 16]) >= 1e-4


def test_backward_matches_finite_differences():
    torch.manual_seed(0)
    s, sc = make_scene(30, W=48, H=48, seed=5, dx=1 / 24)
    means = sc["means"].double().requires_grad_(True)
    cov6, opac, shs = sc["cov6"].double(), sc["opac"].double(), sc["shs"].double()
    binning = RO.bin_np(s, means.detach().numpy(), cov6.numpy())
    G = torch.randn(3, 48, 48, dtype=torch.float64)

    def loss(m):
        return (RO.render_torch(s, binning, m, shs, None, opac, cov6) * G).sum()

    g, = torch.autograd.grad(loss(means), means)
    h = 1e-6
    for (i, j) in [(0, 0), (3, 1), (7, 2), (11, 0), (20, 1)]:
        mp, mm = means.detach().clone(), means.detach().clone()
        mp[i, j] += h
        mm[i, j] -= h
        fd = (loss(mp) - loss(mm)) / (2 * h)
        assert abs(float(fd) - float(g[i, j])) <= 2e-4 * max(1.0, abs(float(fd))), (i, j, float(fd), float(g[i, j]))
