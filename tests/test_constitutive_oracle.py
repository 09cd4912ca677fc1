"""CPU: the constitutive oracle against golden vectors produced by the reference's own classes."""
import os

import numpy as np
import pytest
import torch

from oracle import constitutive_oracle as CO

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MATERIALS = ["jelly", "plasticine", "sand"]


def load_gold(name):
    z = np.load(os.path.join(GOLD, f"material_{name}.npz"))
    return {k: torch.from_numpy(z[k]) for k in z.files}


def weights_and_lora(g, tag):
    p = f"param_{tag}."
    W = [g[p + "layers.0.fc.weight"], g[p + "layers.1.fc.weight"], g[p + "final_layer.fc.weight"]]
    s = float(g["lora_scaling"])
    lora = [(g[p + f"{l}.fc.lora_A"], g[p + f"{l}.fc.lora_B"], s) for l in ("layers.0", "layers.1", "final_layer")]
    return W, lora


@pytest.mark.parametrize("name", MATERIALS)
def test_forward_matches_reference_classes(name):
    g = load_gold(name)
    F = g["F"]
    W, lora = weights_and_lora(g, "elastic")
    out = CO.elasticity(F, W, lora)
    assert torch.allclose(out, g["out_elastic"], rtol=1e-10, atol=1e-10)
    # merged weights (eval mode) == unmerged up to rounding
    out_m = CO.elasticity(F, CO.merge_lora(W, lora))
    assert torch.allclose(out_m, g["out_elastic_merged"], rtol=1e-9, atol=1e-10)
    W, lora = weights_and_lora(g, "plastic")
    out = CO.plasticity(F, W, float(g["alpha"]), lora)
    assert torch.allclose(out, g["out_plastic"], rtol=1e-10, atol=1e-12)


@pytest.mark.parametrize("name", MATERIALS)
def test_backward_matches_reference_autograd(name):
    g = load_gold(name)
    for tag, fn in (("elastic", lambda F, W, l: CO.elasticity(F, W, l)),
                    ("plastic", lambda F, W, l: CO.plasticity(F, W, float(g["alpha"]), l))):
        W, lora = weights_and_lora(g, tag)
        F = g["F"].clone().requires_grad_(True)
        lora = [(A.clone().requires_grad_(True), B.clone().requires_grad_(True), s) for A, B, s in lora]
        y = fn(F, W, lora)
        leaves = [F] + [t for A, B, _ in lora for t in (A, B)]
        grads = torch.autograd.grad(y, leaves, g[f"gout_{tag}"])
        sel = slice(8, None)  # rows 0..7 are F ~ I: SVD adjoint ill-conditioned
        assert torch.allclose(grads[0][sel], g[f"gF_{tag}"][sel], rtol=1e-7, atol=1e-9)
        for i, l in enumerate(("layers.0", "layers.1", "final_layer")):
            assert torch.allclose(grads[1 + 2 * i], g[f"grad_{tag}.{l}.fc.lora_A"], rtol=1e-7, atol=1e-9)
            assert torch.allclose(grads[2 + 2 * i], g[f"grad_{tag}.{l}.fc.lora_B"], rtol=1e-7, atol=1e-9)


def test_svd_signfix_convention():
    torch.manual_seed(0)
    F = torch.randn(200, 3, 3, dtype=torch.float64)
    F[:50] = F[:50] * torch.tensor([1.0, 1.0, -1.0], dtype=torch.float64)  # some det < 0
    U, s, Vh = CO.svd_signfix(F)
    assert torch.all(torch.linalg.det(U) > 0) and torch.all(torch.linalg.det(Vh) > 0)
    assert torch.allclose(U @ torch.diag_embed(s) @ Vh, F, atol=1e-12)
    assert torch.all((s[:, 2] < 0) == (torch.linalg.det(F) < 0))


def test_polar_newton_equals_svd_rotation():
    torch.manual_seed(1)
    F = torch.eye(3, dtype=torch.float64) + 0.2 * torch.randn(100, 3, 3, dtype=torch.float64)
    U, s, Vh = CO.svd_signfix(F)
    assert torch.allclose(CO.polar_newton(F), U @ Vh, atol=1e-10)
