"""Fused frame engine: S MPM substeps (elasticity -> sim -> plasticity), the
particle->Gaussian binding, covariance deformation, rasterization and the complete
reverse pass, driven as a handful of CUDA-graph launches instead of autograd nodes.

It computes exactly what the drop-in classes compute for one rendered frame of
finetune.py's loop (experiments/finetune.py:360-414) -- the per-op modules in
neuma_b200.sim / .material / .glue remain the reference-shaped API; this engine is
the same kernels with the Python/autograd overhead removed:

  * per-substep state is kept compactly: x, v, C, F, stress, F_trial = 168 B per
    particle per substep (the reference's MPMCacheDiffSim pins 528 B,
    modules/nclaw/sim/interface.py:108-123);
  * the S-substep forward and the S-substep backward are each captured once into a
    CUDA graph (static buffers) and replayed;
  * the MPM substeps run the rollout path of the library (csrc/mpm_rollout.cu): particles
    are put in cell (Morton) order at load_state(), every substep keeps its own packed grid
    so the backward reads the forward's grid instead of re-running P2G.  The saved grids are
    COMPACT: a slot holds only the bounding box of the occupied nodes (device-side box, updated
    by G2P for the next substep), `cap_nodes` float4 nodes for the grid and as many for the
    adjoint scatter target -- at 100 k / 128^3 that is 32 B x ~0.2 M nodes = 6 MB per substep
    instead of 67 MB dense (the reference pins 52.8 MB of particle state per substep,
    modules/nclaw/sim/interface.py:108-123).  Everything this class returns or accepts is in
    the CALLER's particle order;
  * LoRA is handled by the caller: the engine takes merged weights and returns
    dL/dW_eff (see neuma_b200.material for the un-merge).
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch

from . import ops
from ._lib import RasterSettingsC, NeumaError


class FusedFrame:
    def __init__(self, n: int, cconst, statics, We: Sequence[torch.Tensor], Wp: Sequence[torch.Tensor],
                 alpha: float, substeps: int, device, bind=None, gaussians=None,
                 settings: Optional[RasterSettingsC] = None, center=None, size=None,
                 elastic_kind: int = ops.KIND_ELASTIC, use_graphs: bool = True, sort: bool = True,
                 compact: bool = True, grid_margin: int = 6, cap_nodes: Optional[int] = None, fuse: bool = True):
        """compact: saved grids hold the node bounding box only; their capacity is fixed at the first
        load_state() from that state's box widened by `grid_margin` nodes on every side (room for
        the cloud to move / deform over the frames this engine will see), or given as cap_nodes.
        A later box that does not fit is reported by check() -- nothing is written out of bounds."""
        self.n, self.S, self.dev = int(n), int(substeps), device
        self.c = cconst
        self._statics_in = tuple(t.contiguous() for t in statics)
        # statics in the engine's (cell-sorted) particle order; refreshed by load_state()
        self.vol, self.rho, self.clip, self.enabled = [t.clone() for t in self._statics_in]
        self.sort = bool(sort)
        self.fuse = bool(fuse)       # P2G / adjoint scatter inside the constitutive kernels' epilogues
        self.perm = torch.arange(n, device=device, dtype=torch.int32)   # sorted slot -> caller index
        self.inv = self.perm.clone()                                    # caller index -> sorted slot
        self.We = [w.detach().clone().contiguous() for w in We]     # engine-owned static buffers (set_weights)
        self.Wp = [w.detach().clone().contiguous() for w in Wp]
        # tensor-core operand images of the weights, refreshed by set_weights()
        self.We_st = ops.constitutive_stage(*self.We)
        self.Wp_st = ops.constitutive_stage(*self.Wp)
        self.alpha = float(alpha)
        self.kind_e = elastic_kind
        self.bind, self.gs, self.settings = bind, gaussians, settings
        f = dict(device=device, dtype=torch.float32)
        S = self.S
        # state[s] = inputs of substep s (s = 0..S-1), state[S] = frame output
        self.x = torch.empty(S + 1, n, 3, **f)
        self.v = torch.empty(S + 1, n, 3, **f)
        self.C = torch.empty(S + 1, n, 3, 3, **f)
        self.F = torch.empty(S + 1, n, 3, 3, **f)
        self.stress = torch.empty(S, n, 3, 3, **f)
        self.Ftrial = torch.empty(S, n, 3, 3, **f)
        # adjoint buffers (ping-pong) and weight-gradient accumulators
        self.g = [[torch.zeros(n, 3, **f), torch.zeros(n, 3, **f), torch.zeros(n, 3, 3, **f),
                   torch.zeros(n, 3, 3, **f)] for _ in range(2)]
        self.gstress = torch.empty(n, 3, 3, **f)
        self.gFtrial = torch.empty(n, 3, 3, **f)
        self.gWe = [torch.zeros_like(w) for w in self.We]
        self.gWp = [torch.zeros_like(w) for w in self.Wp]
        # per-CTA partial sums of the weight gradients (no atomics in the 2 x S backward calls)
        self.gWe_part = ops.constitutive_partials(self.x)
        self.gWp_part = ops.constitutive_partials(self.x)
        self.center = torch.zeros(3, **f) if center is None else center.to(**f)
        self.size = torch.ones(3, **f) if size is None else size.to(**f)
        self._center_h = [float(v) for v in self.center.cpu().tolist()]      # kernel parameters
        self._size_h = [float(v) for v in self.size.cpu().tolist()]
        self.use_graphs = use_graphs
        self._side = torch.cuda.Stream(device)
        self._g_fwd = self._g_bwd = None
        self._ctx = ops._ctx(self.x)
        # per-substep packed grids ((mv, m) -> (v, m) in place), adjoint scatter targets, boxes
        self.compact, self.grid_margin = bool(compact), int(grid_margin)
        self.box = torch.zeros(S + 1, 8, device=device, dtype=torch.int32)
        self.grid = self.adj = None
        if not self.compact:
            self._alloc_grids(cconst.nz * cconst.num_grids * cconst.num_grids)
        elif cap_nodes is not None:
            self._alloc_grids(int(cap_nodes))

    def _alloc_grids(self, nodes: int) -> None:
        f = dict(device=self.dev, dtype=torch.float32)
        if self.compact:
            self.c = self.c.variant(compact=1, cap_nodes=int(nodes))
        self.cap_nodes = int(nodes)
        self.grid = torch.empty(self.S, nodes, 4, **f)
        # ONE adjoint scatter target, cleared (over the substep's box) inside the backward itself
        self.adj = torch.empty(nodes, 4, **f)

    def _cap_from_box(self, box_host) -> int:
        lo, hi = box_host[:3], box_host[3:6]
        if any(h < l for l, h in zip(lo, hi)):
            return 4096
        m, G = self.grid_margin, self.c.num_grids
        ext = [min(h + m, G - 1) - max(l - m, 0) + 1 for l, h in zip(lo, hi)]
        ext[0] = min(G + 7, ((ext[0] + 7) // 8 + 1) * 8)      # x rows are widened to multiples of 8
        return int(ext[0] * ext[1] * ext[2])

    def memory_bytes_per_substep(self) -> dict:
        """What one substep of the rollout keeps resident for the backward."""
        particle = 4 * (3 + 3 + 9 + 9 + 9 + 9) * self.n          # x, v, C, F, stress, F_trial
        grid = 0 if self.grid is None else 16 * self.cap_nodes
        return {"particle_state": particle, "grids": grid, "per_particle": (particle + grid) / max(self.n, 1)}

    def check(self) -> None:
        """Raises if a substep's node box did not fit the compact grid slots (that substep was
        skipped on the device).  Synchronises."""
        if self.compact and self.grid is not None and bool(self.box[:, 6].any().item()):
            raise NeumaError("FusedFrame: the particles' node box outgrew the compact grid slots "
                                 f"(cap_nodes={self.cap_nodes}); build the engine with a larger grid_margin / cap_nodes")

    # ------------------------------------------------------------------ sim
    def _sim_forward(self) -> None:
        c = self.c
        main = torch.cuda.current_stream(self.dev)
        ops.mpm_box_clear(c, self.box[0], self.grid[0], None, self.box[1])
        for s in range(self.S):
            if self.fuse:
                # elasticity + P2G in ONE kernel: the particle is scattered by the thread that just
                # computed its stress (csrc/constitutive.cu, FusedP2G)
                ops.constitutive_fwd_p2g(self.kind_e, self.F[s], self.We_st, self.stress[s], c, self.x[s], self.v[s],
                                         self.C[s], self.vol, self.rho, self.enabled, self.grid[s], box=self.box[s])
            else:
                ops.constitutive_fwd_staged(self.kind_e, self.F[s], self.We_st, 0.0, self.stress[s])
                ops.mpm_p2g_to(c, self.x[s], self.v[s], self.C[s], self.stress[s], self.vol, self.rho, self.enabled,
                               self.grid[s], box=self.box[s])
            ops.mpm_grid_op_box(c, self.box[s], self.grid[s], None)
            ops.mpm_g2p_from(c, self.x[s], self.F[s], self.clip, self.enabled, self.grid[s],
                             (self.x[s + 1], self.v[s + 1], self.C[s + 1], self.Ftrial[s]), self.box[s + 1],
                             box=self.box[s])
            if s + 1 < self.S:
                # the box clear of the NEXT substep's grid slot runs on a side branch next to the
                # plasticity kernel (both only need this substep's G2P)
                self._side.wait_stream(main)
                with torch.cuda.stream(self._side):
                    ops.mpm_box_clear(c, self.box[s + 1], self.grid[s + 1], None, self.box[s + 2])
            ops.constitutive_fwd_staged(ops.KIND_PLASTIC, self.Ftrial[s], self.Wp_st, self.alpha, self.F[s + 1])
            if s + 1 < self.S:
                main.wait_stream(self._side)

    def _sim_backward(self) -> None:
        """Consumes self.g[0] = dL/d(state[S]); leaves dL/d(state[0]) in self.g[0]
        and accumulates dL/dW_eff into gWe / gWp."""
        c = self.c
        main = torch.cuda.current_stream(self.dev)
        cur, nxt = 0, 1
        # the ONE adjoint scatter target is cleared over a substep's box inside the backward itself
        ops.mpm_box_clear(c, self.box[self.S - 1], self.adj, None, None)
        for s in reversed(range(self.S)):
            gx2, gv2, gC2, gF_next = self.g[cur]
            gx, gv, gC, gF = self.g[nxt]
            # plasticity: F[s+1] = P(Ftrial[s]); then adj(G2P) scatters dL/dv_node into adj
            if self.fuse:
                ops.constitutive_bwd_scatter(self.Ftrial[s], self.Wp_st, self.alpha, gF_next, self.gFtrial,
                                             self.gWp_part, c, self.x[s], self.F[s], self.clip, self.enabled,
                                             self.v[s + 1], gx2, gv2, gC2, self.adj, box=self.box[s])
            else:
                ops.constitutive_bwd_partials(ops.KIND_PLASTIC, self.Ftrial[s], self.Wp_st, self.alpha, gF_next,
                                              self.gFtrial, self.gWp_part)
                ops.mpm_bwd_scatter(c, self.x[s], self.F[s], self.clip, self.enabled, self.v[s + 1], gx2, gv2, gC2,
                                    self.gFtrial, self.adj, box=self.box[s])
            ops.mpm_bwd_node_gather(c, self.x[s], self.v[s], self.C[s], self.F[s], self.stress[s], self.vol,
                                    self.rho, self.clip, self.enabled, self.v[s + 1], self.C[s + 1], gx2, gv2, gC2,
                                    self.gFtrial, self.grid[s], self.adj, self.box[s], gx, gv, gC, gF, self.gstress)
            if s > 0:
                # adj was last read by the gather above: clear it for substep s - 1 next to the
                # elasticity backward
                self._side.wait_stream(main)
                with torch.cuda.stream(self._side):
                    ops.mpm_box_clear(c, self.box[s - 1], self.adj, None, None)
            # elasticity: stress[s] = E(F[s]);  dL/dF[s] += E'(F[s])^T gstress
            ops.constitutive_bwd_partials(self.kind_e, self.F[s], self.We_st, 0.0, self.gstress, gF,
                                          self.gWe_part, accumulate_gF=True)
            if s > 0:
                main.wait_stream(self._side)
            cur, nxt = nxt, cur
        ops.constitutive_reduce_partials(self.gWe_part, *self.gWe)
        ops.constitutive_reduce_partials(self.gWp_part, *self.gWp)
        if cur != 0:
            for a, b in zip(self.g[0], self.g[1]):
                a.copy_(b)

    def _capture(self, fn, pre=None):
        """`pre` re-establishes what `fn` consumes."""
        if pre is not None:
            pre()
        fn()                                  # warm-up: grows the library workspaces outside capture
        torch.cuda.synchronize(self.dev)
        side = torch.cuda.Stream(self.dev)
        side.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(side):
            if pre is not None:
                pre()
            fn()
        torch.cuda.current_stream(self.dev).wait_stream(side)
        if pre is not None:
            pre()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            fn()
        return graph

    def load_state(self, x, v, C, F) -> None:
        """State at the start of the frame, in the caller's particle order (device or pinned
        host tensors).  Re-sorts the particles by cell and recomputes the node box."""
        dev = self.dev
        x, v, C, F = [t.to(dev, non_blocking=True).contiguous() for t in (x, v, C, F)]
        if self.sort:
            self.perm.copy_(ops.mpm_cell_order(self.c, x))
            self.inv.copy_(ops.invert_perm(self.perm))
            for dst, src in ((self.x[0], x), (self.v[0], v), (self.C[0], C), (self.F[0], F),
                             (self.vol, self._statics_in[0]), (self.rho, self._statics_in[1]),
                             (self.clip, self._statics_in[2]), (self.enabled, self._statics_in[3])):
                ops.gather_rows(src, self.perm, out=dst)
        else:
            self.x[0].copy_(x)
            self.v[0].copy_(v)
            self.C[0].copy_(C)
            self.F[0].copy_(F)
        ops.mpm_box_from_x(self.c, self.x[0], self.enabled, self.box[0])
        if self.grid is None:                       # first state: size the compact grid slots (one host read)
            self._alloc_grids(self._cap_from_box(self.box[0].cpu().tolist()))

    def _to_caller(self, t: torch.Tensor) -> torch.Tensor:
        return ops.gather_rows(t, self.inv) if self.sort else t

    def _to_sorted(self, t: torch.Tensor, out=None) -> torch.Tensor:
        if self.sort:
            return ops.gather_rows(t.contiguous(), self.perm, out=out)
        if out is not None:
            out.copy_(t)
            return out
        return t

    def sim_forward(self, return_state: bool = True):
        if not self.use_graphs:
            self._sim_forward()
        else:
            if self._g_fwd is None:
                self._g_fwd = self._capture(self._sim_forward)   # slot 0 (state, box) is read-only: replay-safe
            self._g_fwd.replay()
        if not return_state:
            return None
        S = self.S
        return tuple(self._to_caller(t[S]) for t in (self.x, self.v, self.C, self.F))

    def sim_backward(self, gx, gv, gC, gF):
        """Reverse pass of the LAST sim_forward() (may be called again: the adjoint scatter target is
        cleared inside the backward).  Gradients in / out in the caller's order."""
        for dst, src in zip(self.g[0], (gx, gv, gC, gF)):
            if src is None:
                dst.zero_()
            else:
                self._to_sorted(src, out=dst)
        return self._run_backward()

    def _run_backward(self, want_g0: bool = True):
        """self.g[0] holds dL/d(state[S]) in the engine's (cell) order."""
        if not self.use_graphs:
            self._sim_backward()
        else:
            if self._g_bwd is None:
                keep = [t.clone() for t in self.g[0]]
                keep_w = [t.clone() for t in self.gWe + self.gWp]
                self._g_bwd = self._capture(self._sim_backward)
                for dst, src in zip(self.g[0], keep):
                    dst.copy_(src)
                for dst, src in zip(self.gWe + self.gWp, keep_w):
                    dst.copy_(src)
            self._g_bwd.replay()
        return [self._to_caller(t) for t in self.g[0]] if want_g0 else None

    def set_weights(self, We=None, Wp=None) -> None:
        """New (LoRA-merged) weights, e.g. after an optimiser step: copied into the engine's static
        buffers and re-staged, so captured graphs stay valid."""
        for dst, src, st in ((self.We, We, self.We_st), (self.Wp, Wp, self.Wp_st)):
            if src is None:
                continue
            for d, s_ in zip(dst, src):
                d.copy_(s_)
            ops.constitutive_stage(*dst, out=st)

    def zero_weight_grads(self) -> None:
        for t in self.gWe + self.gWp:
            t.zero_()

    # ------------------------------------------------------------------ sim -> render
    def render_forward(self, de_x_prev: torch.Tensor, k_prev: torch.Tensor):
        """particles -> Gaussians (denormalise, both bindings, covariance deformation: ONE launch,
        SURVEY 8f.4) + rasterization of the frame's final state."""
        b, gs, S = self.bind, self.gs, self.S
        means3D, cov, _ = ops.particles_to_gaussians(b.rowptr, b.col, b.w, self.inv if self.sort else None,
                                                     self.x[S], self.F[S], self._center_h, self._size_h,
                                                     de_x_prev, k_prev, gs["cov6"])
        image, radii = ops.raster_fwd(self.settings, means3D, gs["shs"], None, gs["opac"], cov)
        self._render_saved = (means3D, cov, radii)
        return image, means3D

    def render_backward(self, dL_dimage: torch.Tensor, accumulate_into_state_grad: bool = False):
        """dL/dimage -> dL/dmeans3D [K,3] (cov3D carries no gradient).  With
        accumulate_into_state_grad the transpose binding adds it straight into the engine's
        dL/dx buffer (cell order), ready for _run_backward()."""
        b, gs = self.bind, self.gs
        means3D, cov, radii = self._render_saved
        g3, _g2, *_ = ops.raster_bwd(self.settings, means3D, gs["shs"], None, gs["opac"], cov, radii, dL_dimage)
        if accumulate_into_state_grad:
            ops.gaussians_to_particles(b.rowptr, b.col, b.w, self.inv if self.sort else None, g3, self._size_h,
                                       self.g[0][0])
        return g3

    # ------------------------------------------------------------------ one frame, forward + backward
    def final_state(self):
        """(x, v, C, F) after the S substeps, caller order (the next frame's load_state input)."""
        S = self.S
        return tuple(self._to_caller(t[S]) for t in (self.x, self.v, self.C, self.F))

    def frame_forward(self, gt: torch.Tensor, de_x_prev, k_prev, gt_ready: Optional[torch.cuda.Event] = None):
        """S substeps + render + l2 loss (finetune.py:360-389).  Returns (loss, means3D): means3D is the
        next frame's k_prev, (final_state()[0] - center) / size its de_x_prev (both enter detached,
        tune/utils.py:441-446).  gt_ready: event after which `gt` is valid (a host->device copy of the
        ground-truth image issued on another stream overlaps the simulation)."""
        self.sim_forward(return_state=False)
        image, means3D = self.render_forward(de_x_prev, k_prev)
        if gt_ready is not None:
            torch.cuda.current_stream(self.dev).wait_event(gt_ready)
        diff = image - gt
        loss = (diff * diff).mean()
        self._dL_dimage = diff * (2.0 / diff.numel())
        return loss, means3D

    def frame_backward(self, g_state_next=None, want_g0: bool = True):
        """Reverse pass of the last frame_forward(): dL/dW_eff accumulate into gWe / gWp; returns
        dL/d(state at frame start) in the caller's order.  g_state_next = dL/d(this frame's final
        state) coming from the frames after it (caller order; entries may be None)."""
        for i, dst in enumerate(self.g[0]):                     # dL/d(state[S]) in the engine's order
            src = None if g_state_next is None else g_state_next[i]
            if src is None:
                dst.zero_()
            else:
                self._to_sorted(src, out=dst)
        self.render_backward(self._dL_dimage, accumulate_into_state_grad=True)
        return self._run_backward(want_g0)

    def frame_fwd_bwd(self, gt: torch.Tensor, de_x_prev, k_prev, g_state_next=None, want_g0: bool = True,
                      gt_ready: Optional[torch.cuda.Event] = None):
        """frame_forward + frame_backward.  Returns (loss, dL/d(state at frame start))."""
        loss, _ = self.frame_forward(gt, de_x_prev, k_prev, gt_ready)
        return loss, self.frame_backward(g_state_next, want_g0)
