"""SVD autograd module: mirror of /root/reference/modules/nclaw/warp/svd.py
(SVDFunction :9-57, batch_svd :59-96, SVD :99-101), backed by the svd3 kernels."""
from __future__ import annotations

import torch.autograd as autograd
import torch.nn as nn
from torch import Tensor

from . import ops


class SVDFunction(autograd.Function):

    @staticmethod
    def forward(ctx, F: Tensor):
        U, sigma, Vh = ops.svd3_fwd(F.detach())
        ctx.save_for_backward(U, sigma, Vh)
        return U, sigma, Vh

    @staticmethod
    def backward(ctx, grad_U, grad_sigma, grad_Vh):
        U, sigma, Vh = ctx.saved_tensors
        return ops.svd3_bwd(U, sigma, Vh, grad_U, grad_sigma, grad_Vh)


class SVD(nn.Module):
    def forward(self, F: Tensor):
        return SVDFunction.apply(F)
