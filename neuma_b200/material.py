"""Neural constitutive laws: mirror of the two classes every reference script
instantiates (finetune.py:28-31, render.py:24-27),

    InvariantFullMetaElasticity   /root/reference/modules/nclaw/material/meta.py:170-221
    InvariantFullMetaPlasticity   meta.py:442-489
    MLPBlock                      meta.py:20-42      (bias-free Linear + exact GELU)
    LinearLoRA                    material/loralib.py:162-224
    init_lora_layers / freeze_all_except_lora / lora_state_dict   meta.py:185-194

with the same constructor config, parameter names (`layers.0.fc.weight`,
`layers.1.fc.weight`, `final_layer.fc.weight`, `...fc.lora_A/lora_B`) and
train()/eval() merge semantics, so the shipped base_models/*.pt and *_lora.pt
checkpoints load unchanged.  forward() is ONE fused CUDA kernel (SVD, invariants,
MLP, symmetrise, polar recombination); backward ONE more.

The LoRA adaptor is folded into the weight on the host, W_eff = W + scaling * B A,
as a differentiable torch expression evaluated on every forward, so
autograd un-merges dL/dW_eff into dL/dA, dL/dB (and dL/dW if it is trainable).
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn as nn
from torch import Tensor

from . import ops
from .svd import SVD


def _cfg(cfg, k, default=None):
    if isinstance(cfg, dict):
        return cfg.get(k, default)
    return getattr(cfg, k, default)


class LinearLoRA(nn.Linear):
    """loralib.py:162-224 (r, lora_alpha, merge_weights; no dropout / fan_in_fan_out
    -- the reference never sets them, loralib.py:63-68)."""

    def __init__(self, in_features: int, out_features: int, r: int = 0, lora_alpha: int = 1,
                 merge_weights: bool = True, bias: bool = False) -> None:
        super().__init__(in_features, out_features, bias=bias)
        self.r = r
        self.lora_alpha = lora_alpha
        self.merged = False
        self.merge_weights = merge_weights
        if r > 0:
            self.lora_A = nn.Parameter(self.weight.new_zeros((r, in_features)))
            self.lora_B = nn.Parameter(self.weight.new_zeros((out_features, r)))
            self.scaling = self.lora_alpha / self.r
            self.weight.requires_grad = False
            nn.init.kaiming_uniform_(self.lora_A, a=math.sqrt(5))
            nn.init.zeros_(self.lora_B)

    def train(self, mode: bool = True):
        super().train(mode)
        if self.r > 0 and self.merge_weights:
            if mode and self.merged:
                self.weight.data -= (self.lora_B @ self.lora_A) * self.scaling
                self.merged = False
            elif not mode and not self.merged:
                self.weight.data += (self.lora_B @ self.lora_A) * self.scaling
                self.merged = True
        return self

    def effective_weight(self) -> Tensor:
        if self.r > 0 and not self.merged:
            # W + scaling * B A (loralib.py:219-222 folded into the weight), ONE differentiable op
            return torch.addmm(self.weight, self.lora_B, self.lora_A, alpha=self.scaling)
        return self.weight

    def forward(self, x: Tensor) -> Tensor:  # only used outside the fused path
        return nn.functional.linear(x, self.effective_weight(), self.bias)


class MLPBlock(nn.Module):
    """meta.py:20-42 restricted to what the shipped configs use: norm=None,
    no_bias=True, nonlinearity in {gelu, None}."""

    def __init__(self, in_planes: int, out_planes: int, no_bias: bool, norm: Optional[str],
                 nonlinearity: Optional[str]) -> None:
        super().__init__()
        if norm is not None or not no_bias:
            raise NotImplementedError("fused constitutive kernel: only norm=None, no_bias=True is built")
        if nonlinearity is not None and nonlinearity.casefold() != "gelu":
            raise NotImplementedError("fused constitutive kernel: only GELU is built")
        self.fc = nn.Linear(in_planes, out_planes, bias=False)
        nn.init.xavier_uniform_(self.fc.weight)

    def effective_weight(self) -> Tensor:
        fc = self.fc
        return fc.effective_weight() if isinstance(fc, LinearLoRA) else fc.weight


class _ConstitutiveFunction(torch.autograd.Function):

    @staticmethod
    def forward(ctx, kind: int, alpha: float, F: Tensor, W0: Tensor, W1: Tensor, W2: Tensor):
        Fd = F.detach().contiguous()
        Ws = [w.detach().contiguous() for w in (W0, W1, W2)]
        out = ops.constitutive_fwd(kind, Fd, *Ws, alpha=alpha)
        ctx.save_for_backward(Fd, *Ws)
        ctx.kind, ctx.alpha = kind, alpha
        return out

    _partials: dict = {}        # per device: the per-CTA weight-gradient partial sums (zeroed by the reduce)

    @staticmethod
    def backward(ctx, gout: Tensor):
        Fd, W0, W1, W2 = ctx.saved_tensors
        want_w = any(ctx.needs_input_grad[3:6])
        if not want_w:
            gF, g0, g1, g2 = ops.constitutive_bwd(ctx.kind, Fd, W0, W1, W2, ctx.alpha, gout.contiguous(),
                                                  want_wgrad=False)
            return None, None, (gF if ctx.needs_input_grad[2] else None), None, None, None
        # weight gradients through the per-CTA partial sums + one reduce (no same-address atomics:
        # 25 % of the kernel otherwise), as FusedFrame does
        part = _ConstitutiveFunction._partials.get(Fd.device)
        if part is None:
            part = _ConstitutiveFunction._partials[Fd.device] = ops.constitutive_partials(Fd)
        staged = ops.constitutive_stage(W0, W1, W2)
        gF = torch.empty_like(Fd)
        ops.constitutive_bwd_partials(ctx.kind, Fd, staged, ctx.alpha, gout.contiguous(), gF, part)
        gW = [torch.zeros_like(w) for w in (W0, W1, W2)]
        ops.constitutive_reduce_partials(part, *gW)
        return None, None, (gF if ctx.needs_input_grad[2] else None), gW[0], gW[1], gW[2]


class _InvariantFullMeta(nn.Module):
    KIND = ops.KIND_ELASTIC

    def __init__(self, cfg) -> None:
        super().__init__()
        self.dim = 3
        self.svd = SVD()
        widths = list(_cfg(cfg, "layer_widths"))
        if widths != [64, 64]:
            raise NotImplementedError("fused constitutive kernel: layer_widths must be [64, 64]")
        self.normalize_input: bool = bool(_cfg(cfg, "normalize_input", True))
        self.layers = nn.ModuleList()
        width = self.dim + self.dim * self.dim + 1
        for next_width in widths:
            self.layers.append(MLPBlock(width, next_width, _cfg(cfg, "no_bias"), _cfg(cfg, "norm"),
                                        _cfg(cfg, "nonlinearity")))
            width = next_width
        self.final_layer = MLPBlock(width, self.dim * self.dim, _cfg(cfg, "no_bias"), None, None)

    # ---- LoRA management (meta.py:185-194, loralib.py:13-81)
    def init_lora_layers(self, r: int, lora_alpha: int = 1):
        self._weff = None
        for block in list(self.layers) + [self.final_layer]:
            lin = block.fc
            if isinstance(lin, LinearLoRA):
                continue
            lora = LinearLoRA(lin.in_features, lin.out_features, r=r, lora_alpha=lora_alpha)
            lora.weight.data = lin.weight.data
            block.fc = lora.to(lin.weight.device)
        print(f"Initialized LoRA layers in {type(self).__name__} with r={r} and alpha={lora_alpha}.")

    def freeze_all_except_lora(self):
        for n, p in self.named_parameters():
            if "lora_" not in n:
                p.requires_grad = False

    def lora_state_dict(self, bias: str = "none") -> Dict[str, Tensor]:
        sd = self.state_dict()
        return {k: sd[k] for k in sd if "lora_" in k}

    # ---- merged weights: the differentiable expression W + scaling * B A is rebuilt on EVERY
    # forward, as the reference does (loralib.py:216-224) -- three tiny matmuls.  (Round 1 cached
    # it per parameter version, which broke two forward+backward rounds without an optimizer step
    # and went stale under weight.data.copy_().)
    # Opt-in (`module.cache_merged = True`): within ONE rollout the 2 x S forward calls see the same
    # parameters, so the merged weights -- and their autograd node -- can be shared: autograd then sums
    # dL/dW_eff over the calls and runs the LoRA un-merge once instead of 2 x S times (15 small launches
    # per call otherwise).  The shared node is dropped as soon as a backward reaches it (tensor hook),
    # so a second forward+backward round without an optimizer step rebuilds it; it is keyed on the
    # parameters' versions.  NOT seen: writes through `.data` between two forwards of the same rollout
    # (call invalidate_merged()); the reference re-merges on every forward, hence off by default.
    cache_merged: bool = False

    def invalidate_merged(self) -> None:
        self._weff = None

    def effective_weights(self):
        blocks = list(self.layers) + [self.final_layer]
        if not self.cache_merged:
            return [b.effective_weight() for b in blocks]
        key = tuple((p.data_ptr(), p._version, p.requires_grad) for p in self.parameters()) + \
              (torch.is_grad_enabled(), tuple(getattr(b.fc, "merged", False) for b in blocks))
        cached = getattr(self, "_weff", None)
        if cached is None or cached[0] != key:
            ws = [b.effective_weight() for b in blocks]

            def _drop(grad, m=self):
                m._weff = None
                return None
            for w in ws:
                if w.requires_grad and not w.is_leaf:
                    w.register_hook(_drop)
            cached = self._weff = (key, ws)
        return cached[1]

    def _kind(self) -> int:
        return self.KIND

    def _alpha(self) -> float:
        return 0.0

    def forward(self, F: Tensor) -> Tensor:
        W0, W1, W2 = self.effective_weights()
        return _ConstitutiveFunction.apply(self._kind(), self._alpha(), F, W0, W1, W2)


class InvariantFullMetaElasticity(_InvariantFullMeta):
    """meta.py:170-221: F -> Kirchhoff stress R sym(MLP(invariants)) F^T."""

    def _kind(self) -> int:
        return ops.KIND_ELASTIC if self.normalize_input else ops.KIND_ELASTIC_RAW


class InvariantFullMetaPlasticity(_InvariantFullMeta):
    """meta.py:442-489: F -> F + alpha R sym(MLP(invariants)) (always normalised inputs)."""

    def __init__(self, cfg) -> None:
        super().__init__(cfg)
        self.alpha: float = float(_cfg(cfg, "alpha"))

    def _kind(self) -> int:
        return ops.KIND_PLASTIC

    def _alpha(self) -> float:
        return self.alpha


class ComposeMaterial(nn.Module):
    """Per-object section dispatch for multi-object scenes: mirror of
    modules/nclaw/material/preset.py:12-27 (used by experiments/inference.py:265-270 and the
    viewer).  `materials[i]` is evaluated on the i-th contiguous slice of `sections[i]`
    particles; empty sections are skipped; results are concatenated in order."""

    def __init__(self, materials, sections) -> None:
        super().__init__()
        self.materials = nn.ModuleList(materials)
        self.sections = sections

    def update_sections(self, sections) -> None:
        self.sections = sections

    def forward(self, F: Tensor) -> Tensor:
        outs = []
        for m, f in zip(self.materials, torch.split(F, list(self.sections), dim=0)):
            if f.numel() == 0:
                continue
            outs.append(m(f.contiguous()))
        return torch.cat(outs, dim=0)
