"""Generate tests/golden/material_*.npz by running the REFERENCE's own Python classes.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden_material.py

What is the reference's and what is not:
  * InvariantFullMetaElasticity / InvariantFullMetaPlasticity / MLPBlock
    (modules/nclaw/material/meta.py) and LinearLoRA (material/loralib.py) are
    imported from /root/reference and executed unmodified, on the shipped
    experiments/base_models/*.pt weights.
  * `warp` and `omegaconf` are not installable here; they are stubbed just enough
    for the imports to succeed, and each model's `self.svd` (a Warp kernel,
    modules/nclaw/warp/svd.py) is replaced by torch.linalg.svd followed by the
    sign fix of svd.py:76-92.
Outputs are computed in float64 (module.double()) so they can serve as ground truth.
"""
import os
import sys
import types
from unittest import mock

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def stub_modules():
    wp = mock.MagicMock(name="warp")
    wp.Tape = type("Tape", (), {})
    wp.kernel = lambda f: f
    wp.struct = lambda c: c
    sys.modules["warp"] = wp
    oc = types.ModuleType("omegaconf")
    oc.DictConfig = dict
    sys.modules["omegaconf"] = oc


class TorchSVD(torch.nn.Module):
    """torch.linalg.svd + batch_svd's sign fix (svd.py:76-92)."""

    def forward(self, F):
        U, s, Vh = torch.linalg.svd(F)
        V = Vh.transpose(1, 2)
        du = (torch.linalg.det(U) < 0)
        dv = (torch.linalg.det(V) < 0)
        su = torch.where(du, -1.0, 1.0).to(F.dtype)
        sv = torch.where(dv, -1.0, 1.0).to(F.dtype)
        one = torch.ones_like(su)
        U = U * torch.stack([one, one, su], -1)[:, None, :]
        V = V * torch.stack([one, one, sv], -1)[:, None, :]
        s = s * torch.stack([one, one, su * sv], -1)
        return U, s, V.transpose(1, 2)


def main():
    stub_modules()
    sys.path.insert(0, REF)
    from modules.nclaw.material.meta import InvariantFullMetaElasticity, InvariantFullMetaPlasticity

    cfg_e = types.SimpleNamespace(layer_widths=[64, 64], norm=None, nonlinearity="gelu", no_bias=True,
                                  normalize_input=True)
    cfg_p = types.SimpleNamespace(layer_widths=[64, 64], norm=None, nonlinearity="gelu", no_bias=True,
                                  normalize_input=True, alpha=0.001)
    for name in ("jelly", "plasticine", "sand"):
        torch.manual_seed(42)
        ck = torch.load(f"{REF}/experiments/base_models/{name}_0300.pt", weights_only=True, map_location="cpu")
        el = InvariantFullMetaElasticity(cfg_e)
        pl = InvariantFullMetaPlasticity(cfg_p)
        el.load_state_dict(ck["elasticity"])
        pl.load_state_dict(ck["plasticity"])
        el.svd, pl.svd = TorchSVD(), TorchSVD()
        # same order as finetune.py:295-313: init LoRA (r=16, alpha=16), freeze the rest
        el.init_lora_layers(r=16, lora_alpha=16)
        pl.init_lora_layers(r=16, lora_alpha=16)
        el.freeze_all_except_lora()
        pl.freeze_all_except_lora()
        # lora_B is zero-initialised; give it values so the adaptor path is exercised
        for m in (el, pl):
            for n_, p_ in m.named_parameters():
                if "lora_B" in n_:
                    p_.data = 0.05 * torch.randn_like(p_)
        el.double(), pl.double()
        el.train(), pl.train()  # unmerged two-GEMM LoRA path (render.py quirk)

        n = 96
        F = torch.eye(3, dtype=torch.float64) + 0.15 * torch.randn(n, 3, 3, dtype=torch.float64)
        F[:8] = torch.eye(3, dtype=torch.float64) + 1e-3 * torch.randn(8, 3, 3, dtype=torch.float64)
        F.requires_grad_(True)
        g_e = torch.randn(n, 3, 3, dtype=torch.float64)
        g_p = torch.randn(n, 3, 3, dtype=torch.float64)
        out = {"F": F.detach().numpy(), "gout_elastic": g_e.numpy(), "gout_plastic": g_p.numpy(),
               "alpha": np.float64(cfg_p.alpha), "lora_scaling": np.float64(16 / 16)}
        for tag, model, g in (("elastic", el, g_e), ("plastic", pl, g_p)):
            y = model(F)
            params = {k: v for k, v in model.named_parameters() if v.requires_grad}
            grads = torch.autograd.grad(y, [F] + list(params.values()), g)
            out[f"out_{tag}"] = y.detach().numpy()
            out[f"gF_{tag}"] = grads[0].numpy()
            for (k, _), gr in zip(params.items(), grads[1:]):
                out[f"grad_{tag}.{k}"] = gr.numpy()
            for k, v in model.state_dict().items():
                out[f"param_{tag}.{k}"] = v.numpy()
            # merged (eval-mode) forward, inference.py order
            model.eval()
            out[f"out_{tag}_merged"] = model(F).detach().numpy()
            model.train()
        np.savez_compressed(os.path.join(OUT, f"material_{name}.npz"), **out)
        print("wrote", f"material_{name}.npz", {k: v.shape for k, v in out.items() if k.startswith('out')})


if __name__ == "__main__":
    main()
