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
                 compact: bool = True, grid_margin: int = 6, cap_nodes: Optional[int] = None, fuse=True,
                 frames: int = 1, views: Optional[Sequence[RasterSettingsC]] = None, loss: str = "l2",
                 force_mask_data: bool = False, save_svd: Optional[bool] = None):
        """frames: rendered frames this engine rolls out before ONE backward over all of them
        (finetune.py:360-414: loss.backward() comes after nsteps = frames x substeps substeps); the state
        of every substep of every frame stays resident.  views: one GaussianRasterizationSettings per
        camera rendered at each frame (`for view in used_views`, finetune.py:377; default [settings]).
        loss: "l2" | "l1" (loss_utils.py:17-24).  force_mask_data: render with colors_precomp = 1 instead
        of the SH colours (tune/utils.py:390-404).
        compact: saved grids hold the node bounding box only; their capacity is fixed at the first
        load_state() from that state's box widened by `grid_margin` nodes on every side (room for
        the cloud to move / deform over the frames this engine will see), or given as cap_nodes.
        A later box that does not fit is reported by check() -- nothing is written out of bounds.
        save_svd: the forward constitutive kernels keep the right singular vectors V of their inputs
        (2 x 36 B per particle and substep) and the backward kernels skip the Jacobi iteration of the SVD
        (neuma_constitutive_svd_cache).  Measured (profiles/r03s_save_svd_ab.txt): backward -0.045 ms,
        forward +0.05 ms per frame at 100 k / S = 20 -- the Jacobi sweeps are not on the exposed path -- so
        it is OFF by default.
        fuse: which MPM stages ride inside the constitutive kernels.  True (default) | "epilogue" = the two
        scatters (P2G after elasticity, the adjoint scatter after the plasticity backward); "full" = also G2P in
        front of plasticity and the adjoint gather in front of the elasticity backward (three kernels per
        substep each way); False = every stage its own kernel.  Measured (profiles/r03_fuse_modes.txt): with
        programmatic dependent launch on, "epilogue" wins in both directions; "full" costs the forward 13 %.
        "backward-full" = epilogue forward + full backward (the best combination before PDL)."""
        self.n, self.S, self.dev = int(n), int(substeps), device
        self.frames = int(frames)
        self.T = self.S * self.frames                   # substeps resident for one backward
        if loss not in ("l1", "l2"):
            raise ValueError("loss must be 'l1' or 'l2'")
        self.loss_kind, self.force_mask = loss, bool(force_mask_data)
        self.views = list(views) if views is not None else ([settings] if settings is not None else [])
        self.c = cconst
        self._statics_in = tuple(t.contiguous() for t in statics)
        # statics in the engine's (cell-sorted) particle order; refreshed by load_state()
        self.vol, self.rho, self.clip, self.enabled = [t.clone() for t in self._statics_in]
        self.sort = bool(sort)
        if fuse not in (True, False, "full", "epilogue", "backward-full"):
            raise ValueError("fuse must be True, False, 'full', 'epilogue' or 'backward-full'")
        self.fuse = fuse                                 # which MPM stages ride inside the constitutive kernels
        self.fuse_fwd = {True: "epilogue", False: False, "backward-full": "epilogue"}.get(fuse, fuse)
        self.fuse_bwd = {True: "epilogue", False: False, "backward-full": "full"}.get(fuse, fuse)
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
        S, T = self.S, self.T
        # state[s] = inputs of substep s (s = 0..T-1), state[(f + 1) S] = output of frame f
        self.x = torch.empty(T + 1, n, 3, **f)
        self.v = torch.empty(T + 1, n, 3, **f)
        self.C = torch.empty(T + 1, n, 3, 3, **f)
        self.F = torch.empty(T + 1, n, 3, 3, **f)
        self.stress = torch.empty(T, n, 3, 3, **f)
        self.Ftrial = torch.empty(T, n, 3, 3, **f)
        self.save_svd = bool(save_svd)
        self.Ve = torch.empty(T, n, 3, 3, **f) if self.save_svd else None     # V of F[s] (elasticity input)
        self.Vp = torch.empty(T, n, 3, 3, **f) if self.save_svd else None     # V of F_trial[s] (plasticity input)
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
        self._g_fwd, self._g_bwd = {}, {}                # CUDA graphs, one per frame segment
        self._ctx = ops._ctx(self.x)
        # per-substep packed grids ((mv, m) -> (v, m) in place), adjoint scatter targets, boxes
        self.compact, self.grid_margin = bool(compact), int(grid_margin)
        self.box = torch.zeros(T + 1, 8, device=device, dtype=torch.int32)
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
        self.grid = torch.empty(self.T, nodes, 4, **f)
        # adjoint scatter target(s), cleared (over the substep's box) inside the backward itself
        self.adj = torch.empty(nodes, 4, **f)
        self.adj2 = torch.empty(nodes, 4, **f) if self.fuse_bwd == "full" else None

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
        particle = 4 * (3 + 3 + 9 + 9 + 9 + 9 + (18 if self.save_svd else 0)) * self.n   # x, v, C, F, stress, F_trial (+ 2 V)
        grid = 0 if self.grid is None else 16 * self.cap_nodes
        return {"particle_state": particle, "grids": grid, "per_particle": (particle + grid) / max(self.n, 1)}

    def check(self) -> None:
        """Raises if a substep's node box did not fit the compact grid slots (that substep was
        skipped on the device).  Synchronises."""
        if self.grid is not None and bool(self.box[:, 6].any().item()):
            raise NeumaError("FusedFrame: the particles' node box outgrew the compact grid slots "
                                 f"(cap_nodes={self.cap_nodes}; build the engine with a larger grid_margin / cap_nodes)"
                                 " or a particle crossed more than one cell in a substep (use fuse='epilogue')")

    # ------------------------------------------------------------------ sim
    def _sim_forward(self, frame: int = 0) -> None:
        c = self.c
        main = torch.cuda.current_stream(self.dev)
        s0, s1 = frame * self.S, (frame + 1) * self.S
        if self.fuse_fwd == "full":
            # three kernels per substep: [elasticity + P2G] [grid_op + prepare the next slot]
            # [G2P + plasticity] (csrc/constitutive.cu FusedP2G / FusedG2P, mpm_rollout.cu grid_op_box_kernel)
            ops.mpm_box_clear(c, self.box[s0], self.grid[s0], None, None)
            for s in range(s0, s1):
                self._svd_out(self.Ve, s)
                ops.constitutive_fwd_p2g(self.kind_e, self.F[s], self.We_st, self.stress[s], c, self.x[s], self.v[s],
                                         self.C[s], self.vol, self.rho, self.enabled, self.grid[s], box=self.box[s])
                ops.mpm_grid_op_box(c, self.box[s], self.grid[s], None,
                                    next_grid=self.grid[s + 1] if s + 1 < s1 else None, box_reset=self.box[s + 1])
                self._svd_out(self.Vp, s)
                ops.constitutive_fwd_g2p(self.Wp_st, self.alpha, self.F[s + 1], c, self.x[s], self.F[s], self.clip,
                                         self.enabled, self.grid[s],
                                         (self.x[s + 1], self.v[s + 1], self.C[s + 1], self.Ftrial[s]),
                                         box_next=self.box[s + 1], box=self.box[s])
            return
        ops.mpm_box_clear(c, self.box[s0], self.grid[s0], None, self.box[s0 + 1])
        for s in range(s0, s1):
            self._svd_out(self.Ve, s)
            if self.fuse_fwd:
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
            if s + 1 < s1:
                # the box clear of the NEXT substep's grid slot runs on a side branch next to the
                # plasticity kernel (both only need this substep's G2P)
                self._side.wait_stream(main)
                with torch.cuda.stream(self._side):
                    ops.mpm_box_clear(c, self.box[s + 1], self.grid[s + 1], None, self.box[s + 2])
            self._svd_out(self.Vp, s)
            ops.constitutive_fwd_staged(ops.KIND_PLASTIC, self.Ftrial[s], self.Wp_st, self.alpha, self.F[s + 1])
            if s + 1 < s1:
                main.wait_stream(self._side)

    def _svd_out(self, buf, s: int) -> None:
        if buf is not None:
            ops.constitutive_svd_cache(self.x, v_out=buf[s])

    def _svd_in(self, buf, s: int) -> None:
        if buf is not None:
            ops.constitutive_svd_cache(self.x, v_in=buf[s])

    def _sim_backward(self, frame: int = 0) -> None:
        """Consumes self.g[0] = dL/d(state after frame `frame`); leaves dL/d(state at its start) in
        self.g[0] and accumulates dL/dW_eff into the partial-sum buffers (reduced by the caller)."""
        c = self.c
        main = torch.cuda.current_stream(self.dev)
        s0, s1 = frame * self.S, (frame + 1) * self.S
        cur, nxt = 0, 1
        full = self.fuse_bwd == "full"
        # the adjoint scatter target is cleared over a substep's box inside the backward itself
        # (full fusion: two targets in turn, so that the clear for substep s - 1 runs beside substep s)
        adjs = (self.adj, self.adj2) if full else (self.adj, self.adj)
        ops.mpm_box_clear(c, self.box[s1 - 1], adjs[(s1 - 1) & 1], None, None)
        for s in reversed(range(s0, s1)):
            gx2, gv2, gC2, gF_next = self.g[cur]
            gx, gv, gC, gF = self.g[nxt]
            adj = adjs[s & 1]
            if full and s > s0:
                self._side.wait_stream(main)
                with torch.cuda.stream(self._side):
                    ops.mpm_box_clear(c, self.box[s - 1], adjs[(s - 1) & 1], None, None)
            # plasticity: F[s+1] = P(Ftrial[s]); then adj(G2P) scatters dL/dv_node into adj
            self._svd_in(self.Vp, s)
            if self.fuse_bwd:
                ops.constitutive_bwd_scatter(self.Ftrial[s], self.Wp_st, self.alpha, gF_next, self.gFtrial,
                                             self.gWp_part, c, self.x[s], self.F[s], self.clip, self.enabled,
                                             self.v[s + 1], gx2, gv2, gC2, adj, box=self.box[s])
            else:
                ops.constitutive_bwd_partials(ops.KIND_PLASTIC, self.Ftrial[s], self.Wp_st, self.alpha, gF_next,
                                              self.gFtrial, self.gWp_part)
                ops.mpm_bwd_scatter(c, self.x[s], self.F[s], self.clip, self.enabled, self.v[s + 1], gx2, gv2, gC2,
                                    self.gFtrial, adj, box=self.box[s])
            if full:
                # [node pass] [adjoint gather + elasticity backward]
                ops.mpm_bwd_node(c, self.grid[s], adj, self.box[s])
                self._svd_in(self.Ve, s)
                ops.constitutive_bwd_gather(self.kind_e, self.We_st, self.gWe_part, c, self.x[s], self.v[s], self.C[s],
                                            self.F[s], self.stress[s], self.vol, self.rho, self.clip, self.enabled,
                                            self.v[s + 1], self.C[s + 1], gx2, gv2, gC2, self.gFtrial, self.grid[s],
                                            adj, gx, gv, gC, gF, self.gstress, box=self.box[s])
                if s > s0:
                    main.wait_stream(self._side)
                cur, nxt = nxt, cur
                continue
            ops.mpm_bwd_node_gather(c, self.x[s], self.v[s], self.C[s], self.F[s], self.stress[s], self.vol,
                                    self.rho, self.clip, self.enabled, self.v[s + 1], self.C[s + 1], gx2, gv2, gC2,
                                    self.gFtrial, self.grid[s], adj, self.box[s], gx, gv, gC, gF, self.gstress)
            if s > s0:
                # adj was last read by the gather above: clear it for substep s - 1 next to the
                # elasticity backward
                self._side.wait_stream(main)
                with torch.cuda.stream(self._side):
                    ops.mpm_box_clear(c, self.box[s - 1], self.adj, None, None)
            # elasticity: stress[s] = E(F[s]);  dL/dF[s] += E'(F[s])^T gstress
            self._svd_in(self.Ve, s)
            ops.constitutive_bwd_partials(self.kind_e, self.F[s], self.We_st, 0.0, self.gstress, gF,
                                          self.gWe_part, accumulate_gF=True)
            if s > s0:
                main.wait_stream(self._side)
            cur, nxt = nxt, cur
        if frame == 0:                       # the rollout's reverse pass ends here: fold the per-CTA partials
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

    def sim_forward(self, return_state: bool = True, frame: int = 0):
        """S substeps of frame `frame` (its input state is state[frame S]: load_state() for frame 0,
        the previous frame's output otherwise)."""
        if not self.use_graphs:
            self._sim_forward(frame)
        else:
            if frame not in self._g_fwd:          # the segment's first slot (state, box) is read-only: replay-safe
                self._g_fwd[frame] = self._capture(lambda: self._sim_forward(frame))
            self._g_fwd[frame].replay()
        if not return_state:
            return None
        return self.final_state(frame)

    def sim_backward(self, gx, gv, gC, gF):
        """Reverse pass of the LAST sim_forward() (may be called again: the adjoint scatter target is
        cleared inside the backward).  Gradients in / out in the caller's order."""
        for dst, src in zip(self.g[0], (gx, gv, gC, gF)):
            if src is None:
                dst.zero_()
            else:
                self._to_sorted(src, out=dst)
        return self._run_backward()

    def _run_backward(self, want_g0: bool = True, frame: int = 0):
        """self.g[0] holds dL/d(state after frame `frame`) in the engine's (cell) order."""
        if not self.use_graphs:
            self._sim_backward(frame)
        else:
            if frame not in self._g_bwd:
                keep = [t.clone() for t in self.g[0]]
                keep_w = [t.clone() for t in self.gWe + self.gWp + [self.gWe_part, self.gWp_part]]
                self._g_bwd[frame] = self._capture(lambda: self._sim_backward(frame))
                for dst, src in zip(self.g[0], keep):
                    dst.copy_(src)
                for dst, src in zip(self.gWe + self.gWp + [self.gWe_part, self.gWp_part], keep_w):
                    dst.copy_(src)
            self._g_bwd[frame].replay()
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
    def _raster(self, settings, means3D, cov):
        gs = self.gs
        if self.force_mask:                 # tune/utils.py:390-404: colors_precomp = ones, no SH
            if getattr(self, "_ones_rgb", None) is None:
                self._ones_rgb = torch.ones(means3D.shape[0], 3, device=self.dev)
            return ops.raster_fwd(settings, means3D, None, self._ones_rgb, gs["opac"], cov)
        return ops.raster_fwd(settings, means3D, gs["shs"], None, gs["opac"], cov)

    def _raster_bwd(self, settings, means3D, cov, radii, dL_dimage):
        gs = self.gs
        if self.force_mask:
            return ops.raster_bwd(settings, means3D, None, self._ones_rgb, gs["opac"], cov, radii, dL_dimage)[0]
        return ops.raster_bwd(settings, means3D, gs["shs"], None, gs["opac"], cov, radii, dL_dimage)[0]

    def gaussians_at(self, de_x_prev: torch.Tensor, k_prev: torch.Tensor, frame: int = 0):
        """particles -> Gaussians at the end of frame `frame` (denormalise, both bindings, covariance
        deformation: ONE launch): means3D [K,3], deformed cov [K,6]."""
        b, gs, s = self.bind, self.gs, (frame + 1) * self.S
        means3D, cov, _ = ops.particles_to_gaussians(b.rowptr, b.col, b.w, self.inv if self.sort else None,
                                                     self.x[s], self.F[s], self._center_h, self._size_h,
                                                     de_x_prev, k_prev, gs["cov6"])
        return means3D, cov

    def render_forward(self, de_x_prev: torch.Tensor, k_prev: torch.Tensor, frame: int = 0, view: int = 0):
        """gaussians_at() + rasterization of one view."""
        means3D, cov = self.gaussians_at(de_x_prev, k_prev, frame)
        image, radii = self._raster(self.views[view], means3D, cov)
        self._render_saved = (means3D, cov, radii, view)
        return image, means3D

    def render_backward(self, dL_dimage: torch.Tensor, accumulate_into_state_grad: bool = False):
        """dL/dimage of the LAST render_forward -> dL/dmeans3D [K,3] (cov3D carries no gradient,
        SURVEY 0.4).  With accumulate_into_state_grad the transpose binding adds it straight into the
        engine's dL/dx buffer (cell order), ready for _run_backward()."""
        b = self.bind
        means3D, cov, radii, view = self._render_saved
        g3 = self._raster_bwd(self.views[view], means3D, cov, radii, dL_dimage)
        if accumulate_into_state_grad:
            ops.gaussians_to_particles(b.rowptr, b.col, b.w, self.inv if self.sort else None, g3, self._size_h,
                                       self.g[0][0])
        return g3

    def _pixel_loss(self, image: torch.Tensor, gt: torch.Tensor, weight: float):
        """weight * l1 | l2 (loss_utils.py:17-24) and its dL/dimage."""
        diff = image - gt
        if self.loss_kind == "l2":
            return weight * (diff * diff).mean(), diff * (2.0 * weight / diff.numel())
        return weight * diff.abs().mean(), torch.sign(diff) * (weight / diff.numel())

    # ------------------------------------------------------------------ one frame, forward + backward
    def final_state(self, frame: int = 0):
        """(x, v, C, F) after frame `frame`, caller order (the next frame's load_state input)."""
        s = (frame + 1) * self.S
        return tuple(self._to_caller(t[s]) for t in (self.x, self.v, self.C, self.F))

    def frame_forward(self, gt: torch.Tensor, de_x_prev, k_prev, gt_ready: Optional[torch.cuda.Event] = None):
        """S substeps + render + pixel loss (finetune.py:360-389).  Returns (loss, means3D): means3D is the
        next frame's k_prev, (final_state()[0] - center) / size its de_x_prev (both enter detached,
        tune/utils.py:441-446).  gt_ready: event after which `gt` is valid (a host->device copy of the
        ground-truth image issued on another stream overlaps the simulation)."""
        self.sim_forward(return_state=False)
        image, means3D = self.render_forward(de_x_prev, k_prev)
        if gt_ready is not None:
            torch.cuda.current_stream(self.dev).wait_event(gt_ready)
        loss, self._dL_dimage = self._pixel_loss(image, gt, 1.0)
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
        """frame_forward + frame_backward.  Returns (loss, dL/d(state at frame start)).
        With CUDA graphs the WHOLE frame is two graph launches: [S substeps, particles -> Gaussians,
        rasterizer forward] and [pixel loss, rasterizer backward, transpose binding, S-substep reverse
        pass] -- the rasterizer has no host round trip, so it captures like everything else; the
        ground-truth image is waited for (gt_ready) between the two."""
        if self.use_graphs and self.frames == 1 and g_state_next is None and len(self.views) == 1:
            return self._frame_graphed(gt, de_x_prev, k_prev, want_g0, gt_ready)
        loss, _ = self.frame_forward(gt, de_x_prev, k_prev, gt_ready)
        return loss, self.frame_backward(g_state_next, want_g0)

    def _frame_a(self) -> None:
        self._sim_forward(0)
        means3D, cov = self.gaussians_at(self._s_de, self._s_k, 0)
        image, radii = self._raster(self.views[0], means3D, cov)
        self._fa = (means3D, cov, radii, image)

    def _frame_b(self) -> None:
        means3D, cov, radii, image = self._fa
        loss, dimg = self._pixel_loss(image, self._s_gt, 1.0)
        for dst in self.g[0]:
            dst.zero_()
        g3 = self._raster_bwd(self.views[0], means3D, cov, radii, dimg)
        b = self.bind
        ops.gaussians_to_particles(b.rowptr, b.col, b.w, self.inv if self.sort else None, g3, self._size_h,
                                   self.g[0][0])
        self._sim_backward(0)
        self._s_loss = loss

    def _frame_graphed(self, gt, de_x_prev, k_prev, want_g0, gt_ready):
        if getattr(self, "_fg", None) is None:
            self._s_gt, self._s_de, self._s_k = torch.empty_like(gt), torch.empty_like(de_x_prev), torch.empty_like(k_prev)
            self._s_gt.copy_(gt)
            self._s_de.copy_(de_x_prev)
            self._s_k.copy_(k_prev)
            keep = [t.clone() for t in self.gWe + self.gWp + [self.gWe_part, self.gWp_part]]
            self._frame_a()                       # eager warm-up: sizes the library's workspaces
            self._frame_b()
            torch.cuda.synchronize(self.dev)
            ga, gb = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            with torch.cuda.graph(ga):
                self._frame_a()
            with torch.cuda.graph(gb, pool=ga.pool()):
                self._frame_b()
            for dst, src in zip(self.gWe + self.gWp + [self.gWe_part, self.gWp_part], keep):
                dst.copy_(src)
            self._fg = (ga, gb)
        self._s_de.copy_(de_x_prev)
        self._s_k.copy_(k_prev)
        self._fg[0].replay()
        if gt_ready is not None:
            torch.cuda.current_stream(self.dev).wait_event(gt_ready)
        self._s_gt.copy_(gt)
        self._fg[1].replay()
        self._render_saved = self._fa[:3] + (0,)
        g0 = [self._to_caller(t) for t in self.g[0]] if want_g0 else None
        return self._s_loss, g0

    # ------------------------------------------------------------------ the whole rollout of an epoch
    def rollout_fwd_bwd(self, gts, weights=None, de_x_prev=None, k_prev=None, want_g0: bool = True):
        """One epoch of finetune_constitutive (finetune.py:360-414) with ONE engine: `frames` frames of S
        substeps each from the state given to load_state(); after every frame the Gaussians are moved
        with their bound particles and rendered from every view, the weighted pixel losses summed;
        then ONE reverse pass over all frames x substeps.

          gts[f][v]   ground-truth image [3,H,W] (device) of frame f, view v; gts[f] = None skips the
                      frame (`exclude_steps`, finetune.py:369-371): no loss, and the binding reference
                      (de_x_prev, g_prev) is NOT advanced, exactly as the reference's `continue`;
          weights[f]  rollout_decay_rate of frame f (finetune.py:388-389; default 1).
        The loss is a plain weighted sum, so dL/dimage of a view is known as soon as it is rendered:
        its rasterizer backward runs right away and only dL/dmeans3D [K,3] per frame is kept for the
        reverse pass (no rasterizer state is saved).  Returns (loss, dL/d(state0) in the caller's
        order or None); dL/dW_eff are ADDED into gWe / gWp."""
        assert len(gts) == self.frames, "one entry of gts per frame"
        b = self.bind
        dev = self.dev
        if de_x_prev is None:
            x0 = self._to_caller(self.x[0])
            de_x_prev = (x0 - self.center) / self.size
        if k_prev is None:
            raise ValueError("k_prev (the kernels' positions at the start, gaussians.get_xyz) is required")
        loss = torch.zeros((), device=dev)
        if getattr(self, "_gk", None) is None:
            self._gk = torch.zeros(self.frames, k_prev.shape[0], 3, device=dev)
        rendered = []
        for f in range(self.frames):
            self.sim_forward(return_state=False, frame=f)
            if gts[f] is None:
                continue
            w = 1.0 if weights is None else float(weights[f])
            means3D, cov = self.gaussians_at(de_x_prev, k_prev, f)
            gk = self._gk[f]
            for vi, gt in enumerate(gts[f]):
                image, radii = self._raster(self.views[vi], means3D, cov)
                l, dimg = self._pixel_loss(image, gt, w)
                loss = loss + l
                g3 = self._raster_bwd(self.views[vi], means3D, cov, radii, dimg)
                if vi == 0:
                    gk.copy_(g3)
                else:
                    gk.add_(g3)
            rendered.append(f)
            xs = self._to_caller(self.x[(f + 1) * self.S])
            de_x_prev = (xs - self.center) / self.size
            k_prev = means3D
        # ---- reverse pass
        for dst in self.g[0]:
            dst.zero_()
        for f in reversed(range(self.frames)):
            if f in rendered:
                ops.gaussians_to_particles(b.rowptr, b.col, b.w, self.inv if self.sort else None, self._gk[f],
                                           self._size_h, self.g[0][0])
            g0 = self._run_backward(want_g0 and f == 0, frame=f)
        return loss, g0
