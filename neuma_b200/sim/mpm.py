"""Host-side mirror of the reference's MPM model objects, backed by libneuma_b200.so.

Same class names, methods and argument meaning as
/root/reference/modules/nclaw/sim/mpm.py, so finetune.py / render.py run with an
import-path change only:

    MPMStatics              mpm.py:14-72     vol / rho / clip_bound / enabled per particle
    MPMParticleData         mpm.py:75-128    x, v, C, F, stress
    MPMConstant             mpm.py:158-167
    MPMState                mpm.py:170-242   from_torch / to_torch (zero-copy)
    MPMModel                mpm.py:245-319   forward / backward (fused CUDA, no Warp tape)
    MPMModelBuilder         mpm.py:501-550
    MPMInitData             mpm.py:553-688
    MPMStateInitializer     mpm.py:691-735
    MPMStaticsInitializer   mpm.py:738-776

Differences by design: arrays are plain torch CUDA tensors (the reference aliases
torch <-> warp zero-copy, mpm.py:190-242); there is no tape -- `MPMModel.backward`
calls the hand-written adjoint kernels; the grid lives in the library's context.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from pathlib import Path
from typing import Any, Optional, Sequence, Union

import numpy as np
import torch
from torch import Tensor

from .. import ops


def _torch_device(device) -> torch.device:
    if device is None:
        return torch.device("cuda", torch.cuda.current_device())
    if isinstance(device, torch.device):
        return device
    return torch.device(str(device).replace("cuda:", "cuda:") if isinstance(device, str) else device)


class MPMStatics(object):
    """mpm.py:14-72."""

    def init(self, shape: Union[Sequence[int], int], device=None) -> None:
        dev = _torch_device(device)
        self.vol = torch.zeros(shape, dtype=torch.float32, device=dev)
        self.rho = torch.zeros(shape, dtype=torch.float32, device=dev)
        self.clip_bound = torch.zeros(shape, dtype=torch.float32, device=dev)
        self.enabled = torch.zeros(shape, dtype=torch.int32, device=dev)

    @staticmethod
    def _fill(arr: Tensor, sections, values) -> None:
        offset = 0
        for section, value in zip(sections, values):
            arr[offset:offset + section] = value
            offset += section

    def update_vol(self, sections: list, vols: list) -> None:
        self._fill(self.vol, sections, vols)

    def update_rho(self, sections: list, rhos: list) -> None:
        self._fill(self.rho, sections, rhos)

    def update_clip_bound(self, sections: list, clip_bounds: list) -> None:
        self._fill(self.clip_bound, sections, clip_bounds)

    def update_enabled(self, sections: list, spans: list, step: int = 0) -> None:
        self._fill(self.enabled, sections, [1 if (span[0] <= step < span[1]) else 0 for span in spans])


class MPMParticleData(object):
    """mpm.py:75-128 (F initialised to the identity, the rest to zero)."""

    def init(self, shape, device=None, requires_grad: bool = False) -> None:
        dev = _torch_device(device)
        n = int(shape)
        self.x = torch.zeros(n, 3, dtype=torch.float32, device=dev)
        self.v = torch.zeros(n, 3, dtype=torch.float32, device=dev)
        self.C = torch.zeros(n, 3, 3, dtype=torch.float32, device=dev)
        self.F = torch.eye(3, dtype=torch.float32, device=dev).repeat(n, 1, 1)
        self.stress = torch.zeros(n, 3, 3, dtype=torch.float32, device=dev)
        self.requires_grad = requires_grad
        self.grad = dict(x=None, v=None, C=None, F=None, stress=None)

    def clear(self) -> None:
        self.x.zero_(); self.v.zero_(); self.C.zero_(); self.stress.zero_()
        self.F.copy_(torch.eye(3, device=self.F.device).expand_as(self.F))

    def zero_grad(self) -> None:
        for k, g in self.grad.items():
            if g is not None:
                g.zero_()


@dataclass
class MPMConstant(object):
    """mpm.py:158-167."""
    num_grids: int = 0
    dt: float = 0.0
    bound: int = 0
    gravity: Any = (0.0, 0.0, 0.0)
    dx: float = 0.0
    inv_dx: float = 0.0
    eps: float = 0.0


class MPMState(object):
    """mpm.py:170-242."""

    def __init__(self, shape: int, device=None, requires_grad: bool = False) -> None:
        self.shape = shape
        self.device = _torch_device(device)
        self.requires_grad = requires_grad
        particle = MPMParticleData()
        particle.init(shape, self.device, requires_grad)
        self.particle = particle

    def zero_grad(self) -> None:
        self.particle.zero_grad()

    def clear(self) -> None:
        self.particle.clear()

    def to_torch(self):
        p = self.particle
        rg = p.requires_grad
        return tuple(t.detach().requires_grad_(rg) for t in (p.x, p.v, p.C, p.F, p.stress))

    def to_torch_grad(self):
        g = self.particle.grad
        return g["x"], g["v"], g["C"], g["F"], g["stress"]

    def from_torch(self, x: Optional[Tensor] = None, v: Optional[Tensor] = None, C: Optional[Tensor] = None,
                   F: Optional[Tensor] = None, stress: Optional[Tensor] = None) -> None:
        p = self.particle
        if x is not None:
            p.x = x.detach().contiguous()
        if v is not None:
            p.v = v.detach().contiguous()
        if C is not None:
            p.C = C.detach().contiguous()
        if F is not None:
            p.F = F.detach().contiguous()
        if stress is not None:
            p.stress = stress.detach().contiguous()

    def from_torch_grad(self, grad_x=None, grad_v=None, grad_C=None, grad_F=None, grad_stress=None) -> None:
        g = self.particle.grad
        for k, t in (("x", grad_x), ("v", grad_v), ("C", grad_C), ("F", grad_F), ("stress", grad_stress)):
            if t is not None:
                g[k] = t.contiguous()


class MPMModel(object):
    """mpm.py:245-319.  forward = clear + p2g + grid_op + g2p; backward = recompute
    p2g + grid_op, then adj(g2p), adj(grid_op), adj(p2g) -- each ONE fused call."""

    ConstantType = MPMConstant
    StaticsType = MPMStatics
    StateType = MPMState

    def __init__(self, constant: MPMConstant, device=None, requires_grad: bool = False, bc: str = "noslip") -> None:
        self.constant = constant
        self.device = _torch_device(device)
        self.requires_grad = requires_grad
        self.bc = bc

    # the reference swaps model.grid_op between two kernels (mpm.py:545-550)
    def _cconst(self):
        c = self.constant
        return ops.mpm_const(c.num_grids, c.dt, c.bound, tuple(float(g) for g in c.gravity), c.eps, self.bc)

    def state(self, shape: Any, requires_grad: Optional[bool] = None) -> MPMState:
        if requires_grad is None:
            requires_grad = self.requires_grad
        return self.StateType(shape=shape, device=self.device, requires_grad=requires_grad)

    def statics(self, shape: Any) -> MPMStatics:
        statics = self.StaticsType()
        statics.init(shape=shape, device=self.device)
        return statics

    def forward(self, statics: MPMStatics, state_curr: MPMState, state_next: MPMState, tape=None,
                save_grid: bool = False) -> None:
        """save_grid (the differentiable path, MPMSimFunction): the substep runs on the ROLLOUT kernels --
        its (v, m) grid is kept in a compact, box-sized slot owned by `state_curr`, so that backward()
        reads it instead of re-running P2G + grid_op (the reference recomputes them, mpm.py:312-315),
        clears touch the occupied box only, and the node box of the NEXT substep comes out of G2P."""
        pc, pn = state_curr.particle, state_next.particle
        if not save_grid or state_curr is state_next:
            ops.mpm_substep_fwd(self._cconst(), pc.x, pc.v, pc.C, pc.F, pc.stress, statics.vol, statics.rho,
                                statics.clip_bound, statics.enabled, out=(pn.x, pn.v, pn.C, pn.F))
            state_curr.grid = None
            return
        dev = pc.x.device
        if getattr(state_curr, "box", None) is None:
            state_curr.box = torch.zeros(2, 8, dtype=torch.int32, device=dev)
        box = state_curr.box
        # the node box of x: handed over by the previous substep's G2P if x is its output, untouched
        hint = getattr(self, "_box_hint", None)
        if hint is not None and hint[0] == pc.x.data_ptr() and hint[1] == pc.x._version:
            box[0].copy_(hint[2])
        else:
            ops.mpm_box_from_x(self._cconst(), pc.x, statics.enabled, box[0])
        c = self._cconst_compact(box[0])
        if getattr(state_curr, "grid", None) is None or state_curr.grid.shape[0] != c.cap_nodes:
            state_curr.grid = torch.empty(c.cap_nodes, 4, dtype=torch.float32, device=dev)
        ops.mpm_rollout_fwd(c, pc.x, pc.v, pc.C, pc.F, pc.stress, statics.vol, statics.rho, statics.clip_bound,
                            statics.enabled, (pn.x, pn.v, pn.C, pn.F), state_curr.grid, None, box[0], box[1])
        self._box_hint = (pn.x.data_ptr(), pn.x._version, box[1])

    def _cconst_compact(self, box0: Tensor):
        """the C constant with compact grid slots; their capacity is fixed from the first box seen
        (one host read), widened by 6 nodes per side like FusedFrame"""
        cc = getattr(self, "_cc_compact", None)
        if cc is None:
            b = box0.cpu().tolist()
            G, m = self.constant.num_grids, 6
            if any(h < l for l, h in zip(b[:3], b[3:6])):
                cap = 4096
            else:
                ext = [min(h + m, G - 1) - max(l - m, 0) + 1 for l, h in zip(b[:3], b[3:6])]
                ext[0] = min(G + 7, ((ext[0] + 7) // 8 + 1) * 8)
                cap = int(ext[0] * ext[1] * ext[2])
            c = self.constant
            cc = self._cc_compact = ops.mpm_const(c.num_grids, c.dt, c.bound, tuple(float(g) for g in c.gravity),
                                                  c.eps, self.bc, cap_nodes=cap)
            self._adj = torch.empty(cap, 4, dtype=torch.float32, device=box0.device)
        return cc

    def check(self, states) -> None:
        """Raises if a substep's node box outgrew the compact grid slots (that substep was skipped)."""
        for st in states:
            if st is not None and getattr(st, "box", None) is not None and int(st.box[0, 6]) != 0:
                raise ops.NeumaError("MPMModel: the particles' node box outgrew the compact grid slot")

    def backward(self, statics: MPMStatics, state_curr: MPMState, state_next: MPMState, tape=None,
                 need: Sequence[bool] = (True,) * 5) -> None:
        pc, pn = state_curr.particle, state_next.particle
        g = pn.grad
        zeros = lambda t: torch.zeros_like(t)
        gx2 = g["x"] if g["x"] is not None else zeros(pn.x)
        gv2 = g["v"] if g["v"] is not None else zeros(pn.v)
        gC2 = g["C"] if g["C"] is not None else zeros(pn.C)
        gF2 = g["F"] if g["F"] is not None else zeros(pn.F)
        if getattr(state_curr, "grid", None) is not None:
            # saved-grid path: clear the adjoint scatter target over the box, scatter, node pass, gather
            c = self._cc_compact
            outs = [torch.empty_like(t) if nd else None for t, nd in zip((pc.x, pc.v, pc.C, pc.F, pc.stress), need)]
            ops.mpm_box_clear(c, state_curr.box[0], self._adj, None, None)
            ops.mpm_rollout_bwd(c, pc.x, pc.v, pc.C, pc.F, pc.stress, statics.vol, statics.rho, statics.clip_bound,
                                statics.enabled, pn.v, pn.C, gx2.contiguous(), gv2.contiguous(), gC2.contiguous(),
                                gF2.contiguous(), state_curr.grid, self._adj, state_curr.box[0], *outs)
        else:
            outs = ops.mpm_substep_bwd(self._cconst(), pc.x, pc.v, pc.C, pc.F, pc.stress, statics.vol, statics.rho,
                                       statics.clip_bound, statics.enabled, pn.v, gx2, gv2, gC2, gF2, need=need)
        pc.grad = dict(zip(("x", "v", "C", "F", "stress"), outs))


class MPMModelBuilder(object):
    """mpm.py:501-550.  `cfg` is anything with attribute or key access (OmegaConf
    DictConfig, SimpleNamespace, dict)."""

    StateType = MPMState
    ConstantType = MPMConstant
    ModelType = MPMModel

    def __init__(self) -> None:
        self.config = {k: None for k in ("num_grids", "dt", "bound", "gravity", "dx", "inv_dx", "eps")}

    @staticmethod
    def _get(cfg, k):
        return cfg[k] if isinstance(cfg, dict) else getattr(cfg, k)

    def parse_cfg(self, cfg) -> "MPMModelBuilder":
        num_grids = int(self._get(cfg, "num_grids"))
        self.config["num_grids"] = num_grids
        self.config["dt"] = float(self._get(cfg, "dt"))
        self.config["bound"] = int(self._get(cfg, "bound"))
        self.config["gravity"] = np.array(list(self._get(cfg, "gravity")), dtype=np.float32)
        self.config["dx"] = 1 / num_grids
        self.config["inv_dx"] = float(num_grids)
        self.config["bc"] = str(self._get(cfg, "bc"))
        self.config["eps"] = float(self._get(cfg, "eps"))
        return self

    @property
    def ready(self) -> bool:
        return all(v is not None for v in self.config.values())

    def build_constant(self) -> MPMConstant:
        c = self.config
        return MPMConstant(num_grids=c["num_grids"], dt=c["dt"], bound=c["bound"], gravity=c["gravity"],
                           dx=c["dx"], inv_dx=c["inv_dx"], eps=c["eps"])

    def finalize(self, device=None, requires_grad: bool = False) -> MPMModel:
        if not self.ready:
            raise RuntimeError(f"config uninitialized: {self.config}")
        if self.config["bc"] not in ("freeslip", "noslip"):
            raise ValueError("invalid boundary condition: {}".format(self.config["bc"]))
        return self.ModelType(self.build_constant(), device, requires_grad, bc=self.config["bc"])


@dataclass
class MPMInitData(object):
    """mpm.py:553-688 (fields and velocity setters; point clouds come from a
    precomputed <name>.npz -- mesh loading via trimesh is offline asset prep and
    out of scope)."""

    rho: float
    clip_bound: float
    span: tuple
    num_particles: int
    vol: float
    pos: np.ndarray
    lin_vel: np.ndarray = field(default_factory=lambda: np.zeros(3))
    ang_vel: np.ndarray = field(default_factory=lambda: np.zeros(3))
    center: Optional[np.ndarray] = None
    ind_vel: Optional[np.ndarray] = None
    bounds: Optional[np.ndarray] = None
    size: Optional[np.ndarray] = None

    def __post_init__(self) -> None:
        if self.center is None:
            self.center = self.pos.mean(0)

    @staticmethod
    def alignment(min_bound_1, max_bound_1, min_bound_2, max_bound_2):
        center_1 = (min_bound_1 + max_bound_1) / 2
        center_2 = (min_bound_2 + max_bound_2) / 2
        scale_factor = (max_bound_2 - min_bound_2) / (max_bound_1 - min_bound_1)
        translation = center_2 - center_1 * scale_factor
        return scale_factor, translation

    @classmethod
    def get_pcd(cls, name: str, asset_root, sort=None, ori_bounds=None, sim_bounds=None) -> dict:
        assert ori_bounds is not None, "ori_bounds must be provided for pcd shape."
        assert sim_bounds is not None, "sim_bounds must be provided for pcd shape."
        ori_bounds, sim_bounds = np.array(ori_bounds), np.array(sim_bounds)
        path = Path(asset_root) / f"{name}.npz"
        if not path.is_file():
            raise FileNotFoundError(f"{path}: only precomputed .npz point clouds are supported here")
        file = np.load(path)
        p_x, vol = file["p_x"], file["vol"]
        size, center = cls.alignment(ori_bounds[0], ori_bounds[1], sim_bounds[0], sim_bounds[1])
        vol = vol * np.prod(size)
        p_x = np.ascontiguousarray((p_x * size + center).reshape(-1, 3))
        assert p_x.min() >= 0.0 and p_x.max() <= 1.0
        return dict(num_particles=p_x.shape[0], vol=vol, pos=p_x, center=center, size=size)

    def set_lin_vel(self, value) -> None:
        self.lin_vel = np.array(value)

    def zero_lin_vel(self) -> None:
        self.set_lin_vel(np.zeros_like(self.lin_vel))

    def set_ang_vel(self, value) -> None:
        self.ang_vel = np.array(value)

    def zero_ang_vel(self) -> None:
        self.set_ang_vel(np.zeros_like(self.ang_vel))

    def set_ind_vel(self, ind_vel) -> None:
        self.ind_vel = np.array(ind_vel)


class MPMStateInitializer(object):
    """mpm.py:691-735."""

    def __init__(self, model: MPMModel) -> None:
        self.model = model
        self.groups: list = []

    def add_group(self, group: MPMInitData) -> None:
        self.groups.append(group)

    def finalize(self):
        pos_groups, vel_groups, sections = [], [], []
        for group in self.groups:
            pos = group.pos.copy()
            if group.ind_vel is None:
                vel = group.lin_vel.copy() + np.cross(group.ang_vel.copy(), pos - group.center)
            else:
                vel = group.ind_vel.copy()
            pos_groups.append(pos)
            vel_groups.append(np.broadcast_to(vel, pos.shape))
            sections.append(group.num_particles)
        pos_groups = np.concatenate(pos_groups, axis=0)
        vel_groups = np.concatenate(vel_groups, axis=0)
        state_0 = self.model.state(shape=pos_groups.shape[0], requires_grad=False)
        state_0.particle.x.copy_(torch.from_numpy(pos_groups.astype(np.float32)))
        state_0.particle.v.copy_(torch.from_numpy(vel_groups.astype(np.float32)))
        return state_0, sections


class MPMStaticsInitializer(object):
    """mpm.py:738-776."""

    def __init__(self, model: MPMModel) -> None:
        self.model = model
        self.groups: list = []
        self.sections: list = []
        self.vols: list = []
        self.rhos: list = []
        self.clip_bounds: list = []
        self.spans: list = []

    def update(self, statics: MPMStatics, step: int = 0) -> None:
        statics.update_enabled(self.sections, self.spans, step=step)

    def add_group(self, group: MPMInitData) -> None:
        self.groups.append(group)

    def finalize(self) -> MPMStatics:
        for group in self.groups:
            self.sections.append(group.num_particles)
            self.vols.append(group.vol)
            self.rhos.append(group.rho)
            self.clip_bounds.append(group.clip_bound)
            self.spans.append(group.span)
        statics = self.model.statics(sum(self.sections))
        statics.update_vol(self.sections, self.vols)
        statics.update_rho(self.sections, self.rhos)
        statics.update_clip_bound(self.sections, self.clip_bounds)
        self.update(statics, step=0)
        return statics
