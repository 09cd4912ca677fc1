"""torch.autograd seam of the MPM substep: mirror of
/root/reference/modules/nclaw/sim/interface.py (MPMSimFunction :12-76, MPMSim :79-93,
MPMDiffSim :96-105, MPMCacheDiffSim :108-123, MPMForwardSim :126-135).
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.autograd as autograd
import torch.nn as nn
from torch import Tensor

from .mpm import MPMModel, MPMState, MPMStatics


class MPMSimFunction(autograd.Function):

    @staticmethod
    def forward(ctx, model: MPMModel, statics: MPMStatics, state_curr: MPMState, state_next: MPMState,
                x: Tensor, v: Tensor, C: Tensor, F: Tensor, stress: Tensor):
        state_curr.from_torch(x=x, v=v, C=C, F=F, stress=stress)
        # the differentiable path keeps the substep's grid for the backward (no P2G recompute)
        model.forward(statics, state_curr, state_next, None, save_grid=True)
        x_next, v_next, C_next, F_next, _ = state_next.to_torch()
        ctx.model = model
        ctx.statics = statics
        ctx.state_curr = state_curr
        ctx.state_next = state_next
        return x_next, v_next, C_next, F_next

    @staticmethod
    def backward(ctx, grad_x_next, grad_v_next, grad_C_next, grad_F_next):
        model, statics = ctx.model, ctx.statics
        state_curr, state_next = ctx.state_curr, ctx.state_next
        state_next.particle.grad = dict(x=None, v=None, C=None, F=None, stress=None)
        state_next.from_torch_grad(grad_x=grad_x_next, grad_v=grad_v_next, grad_C=grad_C_next,
                                   grad_F=grad_F_next)
        # the reference returns None for inputs whose tensor did not require grad (mpm.py:198-204)
        need = tuple(ctx.needs_input_grad[4:9])
        model.backward(statics, state_curr, state_next, None, need=need)
        # nan_to_num_ (interface.py:65-74) happens inside the adjoint kernel
        grad_x, grad_v, grad_C, grad_F, grad_stress = state_curr.to_torch_grad()
        return None, None, None, None, grad_x, grad_v, grad_C, grad_F, grad_stress


class MPMSim(nn.Module):

    def __init__(self, model: MPMModel) -> None:
        super().__init__()
        self.model = model

    def state(self, x, v, C, F, stress, state: Optional[MPMState] = None) -> MPMState:
        if state is None:
            state = self.model.state(x.size(0))
        state.from_torch(x=x, v=v, C=C, F=F, stress=stress)
        return state


class MPMDiffSim(MPMSim):

    def forward(self, statics: MPMStatics, x, v, C, F, stress):
        shape = x.size(0)
        return MPMSimFunction.apply(self.model, statics, self.model.state(shape), self.model.state(shape),
                                    x, v, C, F, stress)


class MPMCacheDiffSim(MPMSim):

    def __init__(self, model: MPMModel, num_steps: int) -> None:
        super().__init__(model)
        self.curr_states = [None for _ in range(num_steps)]
        self.next_states = [None for _ in range(num_steps)]

    def forward(self, statics: MPMStatics, step: int, x, v, C, F, stress):
        shape = x.size(0)
        if self.curr_states[step] is None:
            self.curr_states[step] = self.model.state(shape)
        if self.next_states[step] is None:
            self.next_states[step] = self.model.state(shape)
        return MPMSimFunction.apply(self.model, statics, self.curr_states[step], self.next_states[step],
                                    x, v, C, F, stress)


class MPMForwardSim(MPMSim):

    def forward(self, statics: MPMStatics, state: MPMState):
        self.model.forward(statics, state, state, None)   # curr and next alias (interface.py:133)
        x_next, v_next, C_next, F_next, _ = state.to_torch()
        return x_next, v_next, C_next, F_next
