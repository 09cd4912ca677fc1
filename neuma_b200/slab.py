"""Z-slab domain decomposition of the MPM rollout across the GPUs of one box.

The reference is single-GPU (no torch.distributed anywhere); this is the multi-GPU design
the north_star asks for: particles and grid are sharded by z, one process per GPU, and the
only data-path communication of a substep is ONE summing exchange of the overlapping grid
planes with the two z-neighbours (NCCL send/recv over NVLink), forward and backward:

    forward  substep:  clear, P2G (local) -> exchange-sum (mv, m) planes -> grid_op, G2P (local)
    backward substep:  scatter adj(G2P) (local) -> exchange-sum dL/dv planes -> adj(grid_op), gather

Rank r owns the particles whose stencil base plane  base_z = int(z * G - 0.5)  lies in
[b[r], b[r+1]).  A particle touches planes base_z .. base_z + 2, and is allowed to drift by
`k` planes before it has to change owner, so rank r holds planes
[b[r] - k, b[r+1] + 2 + k).  In the plane-major packed layout (z slowest, include/
neuma_b200.h) the planes it shares with rank r+1 -- [b[r+1] - k, b[r+1] + 2 + k) -- are one
contiguous block: they are sent straight out of the grid, received into a scratch block and
added with neuma_mpm_planes_add.  Both neighbours add the same two fp32 numbers, so both end
up with bit-identical totals.  Slab boundaries are chosen for equal particle counts, not
equal thickness.  Particles change owner between frames (SlabRollout.repartition);
weight gradients are all-reduced once per frame.

`SlabPlan` and `exchange_sum` are pure host / torch.distributed logic and also run on CPU
tensors with the gloo backend (tests/test_slab_cpu.py); `SlabRollout` drives the sm_100a
kernels.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist


class SlabPlan:
    """Slab boundaries `b` (len world + 1) over stencil-base planes 0 .. G-1."""

    def __init__(self, G: int, bounds: Sequence[int], k: int = 1):
        self.G, self.k = int(G), int(k)
        self.b = [int(v) for v in bounds]
        self.world = len(self.b) - 1
        if self.b[0] != 0 or self.b[-1] != self.G or any(b1 <= b0 for b0, b1 in zip(self.b, self.b[1:])):
            raise ValueError(f"bad slab bounds {self.b}")
        if self.world > 1 and min(b1 - b0 for b0, b1 in zip(self.b, self.b[1:])) < self.min_thickness(k):
            raise ValueError(f"slabs thinner than {self.min_thickness(k)} planes: {self.b}")

    @staticmethod
    def min_thickness(k: int) -> int:
        # the overlap with the lower and the upper neighbour ([b-k, b+2+k) each) must not meet
        return 2 + 2 * k

    @staticmethod
    def base_plane(z: torch.Tensor, G: int) -> torch.Tensor:
        """int(z * G - 0.5): truncation toward zero, as mpm.py:337-339."""
        return (z.float() * float(G) - 0.5).to(torch.int64).clamp_(0, G - 1)

    @classmethod
    def from_histogram(cls, hist: Sequence[int], world: int, k: int = 1) -> "SlabPlan":
        """Boundaries with (nearly) equal particle counts per slab, every slab at least
        min_thickness planes thick.  `hist[z]` = number of particles with base plane z."""
        hist = np.asarray(hist, dtype=np.int64)
        G = int(hist.shape[0])
        t = cls.min_thickness(k)
        if world * t > G:
            raise ValueError(f"{world} slabs of >= {t} planes do not fit in {G} planes")
        cum = np.concatenate([[0], np.cumsum(hist)])
        total = int(cum[-1])
        b = [0]
        for r in range(1, world):
            target = total * r / world
            z = int(np.searchsorted(cum, target, side="left"))
            z = max(z, b[-1] + t)                       # this slab thick enough
            z = min(z, G - (world - r) * t)             # room for the slabs above
            b.append(z)
        b.append(G)
        return cls(G, b, k)

    def owner_of(self, base_z: torch.Tensor) -> torch.Tensor:
        edges = torch.tensor(self.b[1:-1], dtype=torch.int64, device=base_z.device)
        return torch.bucketize(base_z.to(torch.int64), edges, right=True)

    def local_planes(self, rank: int) -> Tuple[int, int]:
        """(z_off, nz) of the planes rank `rank` holds."""
        lo = 0 if rank == 0 else self.b[rank] - self.k
        hi = self.G if rank == self.world - 1 else min(self.G, self.b[rank + 1] + 2 + self.k)
        return lo, hi - lo

    def overlap(self, rank: int, up: bool) -> Optional[Tuple[int, int]]:
        """(z_start, nplanes) shared with rank+1 (up) / rank-1 (down); None at the ends."""
        nb = rank + 1 if up else rank - 1
        if nb < 0 or nb >= self.world:
            return None
        edge = self.b[rank + 1] if up else self.b[rank]
        lo, hi = edge - self.k, min(self.G, edge + 2 + self.k)
        return lo, hi - lo

    def tolerated(self, rank: int) -> Tuple[int, int]:
        """base planes [lo, hi) a particle of `rank` may sit in without leaving its local grid."""
        lo = 0 if rank == 0 else self.b[rank] - self.k
        hi = self.G if rank == self.world - 1 else self.b[rank + 1] + self.k
        return lo, hi


def exchange_sum(planes: torch.Tensor, plan: SlabPlan, rank: int, group=None, add=None,
                 scratch: Optional[dict] = None) -> None:
    """ONE summing halo exchange: `planes` is the rank's [nz, ...] plane-major grid block; the
    blocks shared with rank-1 / rank+1 are exchanged (send/recv pairs) and the neighbour's
    values added in place.  `add(dst, src)` defaults to dst.add_(src) (CPU / tests); the GPU
    path passes ops.mpm_planes_add."""
    if plan.world == 1:
        return
    z_off, _ = plan.local_planes(rank)
    ops_, recv = [], []
    for up in (False, True):
        ov = plan.overlap(rank, up)
        if ov is None:
            continue
        zs, npl = ov
        view = planes[zs - z_off: zs - z_off + npl]
        key = ("up" if up else "down", tuple(view.shape), view.dtype, view.device)
        buf = None if scratch is None else scratch.get(key)
        if buf is None:
            buf = torch.empty_like(view)
            if scratch is not None:
                scratch[key] = buf
        peer = rank + 1 if up else rank - 1
        gpeer = peer if group is None else dist.get_global_rank(group, peer)
        ops_.append(dist.P2POp(dist.isend, view, gpeer, group))
        ops_.append(dist.P2POp(dist.irecv, buf, gpeer, group))
        recv.append((view, buf))
    for req in dist.batch_isend_irecv(ops_):
        req.wait()
    for view, buf in recv:
        if add is None:
            view.add_(buf)
        else:
            add(view, buf)


class PeerHalo:
    """The summing halo exchange over NVLink PEER MEMORY instead of NCCL send/recv, two of our own
    kernels per neighbour (csrc/mpm_rollout.cu):

      neuma_halo_push     waits until the neighbour has acknowledged the previous exchange, copies my
                          shared planes straight into the neighbour's mailbox (stores to a peer
                          pointer over NVLink / NVSwitch) and publishes the exchange number in the
                          neighbour's data flag (st.release.sys);
      neuma_halo_add_ack  waits for my data flag (ld.acquire.sys), adds my mailbox into my grid and
                          acknowledges in the neighbour's ack flag.

    Mailboxes and flags live in torch symmetric memory (peer-mapped allocations); flags carry
    monotonically increasing exchange numbers kept in device counters, so the kernels take only
    pointers and the whole S-substep rollout INCLUDING its 2 S exchanges is captured into one CUDA
    graph per direction (NCCL point-to-point under capture hangs on this stack; measured)."""

    def __init__(self, G: int, k: int, device, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        from . import ops
        self.ops = ops
        self.group = dist.group.WORLD if group is None else group
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        self.npl, self.G = 2 + 2 * k, int(G)
        self.shape = (2, self.npl, self.G, self.G, 4)          # [from lower | from upper][planes][y][x][4]
        self.box = symm_mem.empty(self.shape, dtype=torch.float32, device=device)
        self.hdl = symm_mem.rendezvous(self.box, self.group)
        # flags written by the neighbours: [data from lower, data from upper, ack from lower, ack from upper]
        self.flags = symm_mem.empty((8,), dtype=torch.int32, device=device)
        self.fhdl = symm_mem.rendezvous(self.flags, self.group)
        self.flags.zero_()
        # local counters: [seq push->lower, seq push->upper, seq recv<-lower, seq recv<-upper, done x4]
        self.cnt = torch.zeros(8, dtype=torch.int32, device=device)
        torch.cuda.synchronize(device)
        self.hdl.barrier(0)                                    # every rank's flags are zero before anyone pushes
        nbs = [nb for nb in (self.rank - 1, self.rank + 1) if 0 <= nb < self.world]
        self.peer_box = {nb: self.hdl.get_buffer(nb, self.shape, torch.float32, 0) for nb in nbs}
        self.peer_flags = {nb: self.fhdl.get_buffer(nb, (8,), torch.int32, 0) for nb in nbs}

    def exchange(self, planes: torch.Tensor, plan: SlabPlan, add=None) -> None:
        ops, z_off = self.ops, plan.local_planes(self.rank)[0]
        todo = []
        for up in (False, True):
            ov = plan.overlap(self.rank, up)
            if ov is None:
                continue
            nb = self.rank + 1 if up else self.rank - 1
            view = planes[ov[0] - z_off: ov[0] - z_off + ov[1]]
            u = 1 if up else 0
            # seen from the neighbour ABOVE me I am its lower side (slot 0), and vice versa
            ops.halo_push(view, self.peer_box[nb][1 - u, : ov[1]], self.peer_flags[nb][1 - u: 2 - u],
                          self.flags[2 + u: 3 + u], self.cnt[u: u + 1], self.cnt[4 + u: 5 + u])
            todo.append((nb, view, u, ov[1]))
        for nb, view, u, npl in todo:
            ops.halo_add_ack(view, self.box[u, :npl], self.flags[u: u + 1], self.peer_flags[nb][3 - u: 4 - u],
                             self.cnt[2 + u: 3 + u], self.cnt[6 + u: 7 + u])


class FusedHalo:
    """The summing halo exchange as ONE kernel per exchange for both neighbours
    (csrc/mpm_rollout.cu::halo_exchange_kernel, neuma_halo_exchange): pushes only the x-y window
    this rank's particles occupy, double-buffered mailboxes in torch symmetric memory
    (peer-mapped over NVLink / NVSwitch), monotonic DATA / ACK flags -- pointers only, so the
    whole rollout including its exchanges replays from a CUDA graph, and the kernel can run on a
    side stream next to the interior P2G / G2P (SlabRollout overlap)."""

    def __init__(self, G: int, k: int, device, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        from . import ops
        self.ops = ops
        self.group = dist.group.WORLD if group is None else group
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        self.npl, self.G = 2 + 2 * k, int(G)
        # [from lower | from upper][slot][planes][y][x][4]
        self.mshape = (2, 2, self.npl, self.G, self.G, 4)
        self.mail = symm_mem.empty(self.mshape, dtype=torch.float32, device=device)
        self.mhdl = symm_mem.rendezvous(self.mail, self.group)
        # ints: [0:2] data flag from lower / upper, [2:4] ack flag from lower / upper,
        #       [8:24] window header from lower (2 slots x 8), [24:40] from upper
        self.ints = symm_mem.empty((48,), dtype=torch.int32, device=device)
        self.ihdl = symm_mem.rendezvous(self.ints, self.group)
        self.ints.zero_()
        self.cnt = torch.zeros(4, dtype=torch.int32, device=device)
        torch.cuda.synchronize(device)
        self.mhdl.barrier(0)                                   # every rank's flags are zero before anyone pushes
        nbs = [nb for nb in (self.rank - 1, self.rank + 1) if 0 <= nb < self.world]
        self.peer_mail = {nb: self.mhdl.get_buffer(nb, self.mshape, torch.float32, 0) for nb in nbs}
        self.peer_ints = {nb: self.ihdl.get_buffer(nb, (48,), torch.int32, 0) for nb in nbs}

    def _side(self, planes: torch.Tensor, plan: "SlabPlan", up: bool):
        ov = plan.overlap(self.rank, up)
        if ov is None:
            return None
        z_off = plan.local_planes(self.rank)[0]
        nb = self.rank + 1 if up else self.rank - 1
        u = 1 if up else 0                 # my slot for data FROM this neighbour
        o = 1 - u                          # seen from that neighbour I am its other side
        pi, mi = self.peer_ints[nb], self.ints
        view = planes[ov[0] - z_off: ov[0] - z_off + ov[1]]
        return self.ops.halo_side(view, self.peer_mail[nb][o], pi[8 + 16 * o: 24 + 16 * o], self.mail[u],
                                  mi[8 + 16 * u: 24 + 16 * u], pi[o: o + 1], mi[u: u + 1], pi[2 + o: 3 + o],
                                  mi[2 + u: 3 + u], ov[1], ov[0])

    def exchange(self, planes: torch.Tensor, plan: "SlabPlan", c, box: torch.Tensor) -> None:
        lo, hi = self._side(planes, plan, False), self._side(planes, plan, True)
        self.ops.halo_exchange(c, lo, hi, box, self.cnt)


def split_on_tile_rounds(nb: int, n: int, sms: int, pipelines: int, tile: int = 128) -> int:
    """Smallest split point >= nb that is a whole number of tile rounds of a persistent constitutive kernel
    (one CTA per SM, `pipelines` tile pipelines of `tile` particles each), clipped to n: launching [0, split)
    and [split, n) separately then costs no more rounds than one launch over [0, n)."""
    rnd = sms * pipelines * tile
    return min(n, -(-nb // rnd) * rnd)


class SlabRollout:
    """S-substep differentiable rollout (elasticity -> MPM -> plasticity) of this rank's slab.

    Same kernels and per-substep saved grids as neuma_b200.pipeline.FusedFrame, plus the halo
    exchange:
      halo="fused" (default)  ONE own kernel per exchange (FusedHalo), windowed to the box the rank's
                              particles occupy.  overlap="split" (default): the particles are kept boundary
                              class first; the boundary range's scatter kernel, then the exchange on a side
                              stream under the interior range's scatter kernel -- split on a whole number of
                              tile rounds of the persistent constitutive kernels (split_fwd / split_bwd).
                              overlap="none": in line.  overlap="scatter" / "full": run-time particle
                              classes (csrc/mpm_common.cuh::particle_selected), kept as measured alternatives;
      halo="peer"             round 1: push + add/ack kernel pair per neighbour (PeerHalo);
      halo="nccl"             send/recv pairs, eager launches.
    The backward runs [plasticity bwd + adjoint scatter] [exchange] [node pass] [adjoint gather + elasticity
    bwd] with two adjoint targets in turn (overlap "none" / "split", fuse=True)."""

    def __init__(self, G: int, dt: float, We, Wp, alpha: float, substeps: int, device, group=None,
                 bound: int = 1, gravity=(0.0, -9.8, 0.0), eps: float = 6e-7, bc: str = "noslip",
                 halo_k: int = 1, elastic_kind: int = 0, halo: str = "fused", use_graphs: Optional[bool] = None,
                 overlap="split", standalone: bool = False, fuse: bool = True, fuse_gather: bool = False):
        from . import ops
        self.ops = ops
        self.G, self.dt, self.S, self.dev, self.group = int(G), float(dt), int(substeps), device, group
        self.cargs = (int(bound), tuple(gravity), float(eps), bc)
        self.k = int(halo_k)
        self.alpha, self.kind_e = float(alpha), elastic_kind
        # standalone: this process runs the whole cloud alone even if a process group exists
        # (bench.py times the 1-GPU frame on rank 0 beside the sharded one)
        self.world = dist.get_world_size(group) if (dist.is_initialized() and not standalone) else 1
        self.rank = dist.get_rank(group) if (dist.is_initialized() and not standalone) else 0
        self.skip_exchange = False          # timing experiments only: results are wrong without the exchange
        self.We = [w.detach().to(device).contiguous() for w in We]
        self.Wp = [w.detach().to(device).contiguous() for w in Wp]
        self.We_st = ops.constitutive_stage(*self.We)
        self.Wp_st = ops.constitutive_stage(*self.Wp)
        self.gWe = [torch.zeros_like(w) for w in self.We]
        self.gWp = [torch.zeros_like(w) for w in self.Wp]
        self.gWe_part = ops.constitutive_partials(self.We[0])
        self.gWp_part = ops.constitutive_partials(self.We[0])
        self.plan: Optional[SlabPlan] = None
        self.n = 0
        self._scratch: dict = {}
        if halo not in ("fused", "peer", "nccl"):
            raise ValueError("halo must be 'fused', 'peer' or 'nccl'")
        self.halo = halo
        self._peer = PeerHalo(G, self.k, device, group) if (halo == "peer" and self.world > 1) else None
        self._fused = FusedHalo(G, self.k, device, group) if (halo == "fused" and self.world > 1) else None
        # overlap of the exchange with particle work (fused halo only):
        #   "scatter": the scatter pass (P2G / adjoint scatter) runs boundary class first, the exchange
        #              goes to a side stream under the interior class's scatter, one join before the
        #              node pass (ONE extra launch per exchange);
        #   "full":    the interior class also runs its node pass and gather (G2P / adjoint gather)
        #              under the exchange (four extra launches per exchange);
        #   "none" / False: everything in line.
        #   "split":   the particles are kept boundary class first (static classes with a margin of k
        #              planes, decided at distribute()); elasticity + P2G of the boundary range, then
        #              the exchange on a side stream under elasticity + P2G of the interior range
        #              (backward: plasticity backward + adjoint scatter likewise).  An interior-range
        #              particle that touches a shared plane after all is counted on the device.
        # Measured (profiles/r02k_slab_scaling.txt, r03_slab_split.txt): "scatter" / "full" cost in extra launches
        # what they hide; "split" pays once its split points sit on whole tile ROUNDS of the persistent
        # constitutive kernels (distribute(): split_fwd / split_bwd) -- 1M particles on 2 GPUs 21.2 -> 20.4 ms
        # per frame against a no-exchange floor of 20.2 -- and is the default.
        overlap = {True: "split", False: "none", None: "none"}.get(overlap, overlap)
        if overlap not in ("none", "scatter", "full", "split"):
            raise ValueError("overlap must be 'none', 'split', 'scatter' or 'full'")
        self.overlap = overlap if self._fused is not None else "none"
        self.fuse = bool(fuse)      # P2G / adjoint scatter inside the constitutive kernels' epilogues
        # the adjoint gather inside the elasticity backward's prologue (+ two adjoint targets in turn): the better
        # backward before programmatic dependent launch was on, the worse one since (profiles/r03_fuse_modes.txt)
        self.fuse_gather = bool(fuse_gather) and self.fuse
        # graphs need capturable communication: peer memory yes, NCCL send/recv no
        self.use_graphs = (halo != "nccl" or self.world == 1) if use_graphs is None else bool(use_graphs)
        if self.use_graphs and halo == "nccl" and self.world > 1:
            raise ValueError("CUDA graphs need halo='fused' or 'peer' (NCCL send/recv is not capturable here)")
        self._graphs: dict = {}
        self._side = torch.cuda.Stream(device)

    # ------------------------------------------------------------------ distribution
    def distribute(self, x, v, C, F, vol, rho, clip, enabled) -> None:
        """Every rank passes the same GLOBAL arrays (any device); the rank keeps its slab's
        particles, in cell order, and remembers their global indices in `self.ids`."""
        ops, G, dev = self.ops, self.G, self.dev
        base = SlabPlan.base_plane(x[:, 2], G)
        hist = torch.bincount(base.cpu(), minlength=G)
        self.plan = SlabPlan.from_histogram(hist.numpy(), self.world, self.k)
        mine = (self.plan.owner_of(base) == self.rank).nonzero(as_tuple=True)[0]
        z_off, nz = self.plan.local_planes(self.rank)
        self.c = ops.mpm_const(G, self.dt, self.cargs[0], self.cargs[1], self.cargs[2], self.cargs[3], z_off, nz)
        # planes below lo_end / from up_start on are shared with the lower / upper neighbour
        ov_lo, ov_up = self.plan.overlap(self.rank, False), self.plan.overlap(self.rank, True)
        lo_end = ov_lo[0] + ov_lo[1] if ov_lo is not None else -(1 << 30)
        up_start = ov_up[0] if ov_up is not None else (1 << 30)
        self.c_bnd = self.c.variant(sel_mode=1, sel_lo_end=lo_end, sel_up_start=up_start)
        self.c_int = self.c.variant(sel_mode=2, sel_lo_end=lo_end, sel_up_start=up_start)
        self.c_chk = self.c.variant(sel_mode=3, sel_lo_end=lo_end, sel_up_start=up_start)
        loc = [t[mine.to(t.device)].to(dev).contiguous() for t in (x, v, C, F, vol, rho, clip, enabled)]
        perm = ops.mpm_cell_order(self.c, loc[0]) if loc[0].shape[0] else torch.zeros(0, dtype=torch.int32, device=dev)
        # boundary class first (cell order inside each class): particles whose stencil can reach a
        # shared plane within the frame -- the exact criterion widened by the drift allowance k
        bz = SlabPlan.base_plane(loc[0][:, 2], G)[perm.long()]
        bnd = (bz < lo_end + self.k) | (bz + 2 >= up_start - self.k)
        order = torch.cat([bnd.nonzero(as_tuple=True)[0], (~bnd).nonzero(as_tuple=True)[0]])
        perm = perm[order].contiguous()
        self.nb = int(bnd.sum().item())
        # overlap="split" launches [0, split) first and [split, n) under the exchange.  The constitutive kernels
        # are persistent (one CTA per SM, 3 forward / 2 backward tile pipelines of 128 particles): a split
        # point on a whole number of tile ROUNDS costs no extra round, any other point costs one
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        n_loc = int(loc[0].shape[0]) if loc[0].dim() else 0
        self.split_fwd = split_on_tile_rounds(self.nb, n_loc, sms, 3)
        self.split_bwd = split_on_tile_rounds(self.nb, n_loc, sms, 2)
        loc = [ops.gather_rows(t, perm) if t.shape[0] else t for t in loc]
        self.ids = mine.to(dev)[perm.long()]
        n = self.n = int(loc[0].shape[0])
        S = self.S
        f = dict(device=dev, dtype=torch.float32)
        self.x = torch.empty(S + 1, n, 3, **f); self.v = torch.empty(S + 1, n, 3, **f)
        self.C = torch.empty(S + 1, n, 3, 3, **f); self.F = torch.empty(S + 1, n, 3, 3, **f)
        self.stress = torch.empty(S, n, 3, 3, **f); self.Ftrial = torch.empty(S, n, 3, 3, **f)
        self.x[0], self.v[0], self.C[0], self.F[0] = loc[:4]
        self.vol, self.rho, self.clip, self.enabled = loc[4:]
        self.grid = torch.zeros(S, nz, G, G, 4, **f)
        self.adj = torch.zeros(nz, G, G, 4, **f)           # adjoint scatter target(s), cleared in the backward
        self.adj2 = torch.zeros(nz, G, G, 4, **f) if (self.fuse_gather and self.overlap in ("none", "split")) else None
        self.box = torch.zeros(S + 1, 8, device=dev, dtype=torch.int32)
        self.g = [[torch.zeros(n, 3, **f), torch.zeros(n, 3, **f), torch.zeros(n, 3, 3, **f),
                   torch.zeros(n, 3, 3, **f)] for _ in range(2)]
        self.gstress = torch.empty(n, 3, 3, **f); self.gFtrial = torch.empty(n, 3, 3, **f)
        ops.mpm_box_from_x(self.c, self.x[0], self.enabled, self.box[0])
        self._scratch = {}
        self._graphs = {}                              # buffers were re-allocated: captured graphs are stale

    def _overlap_views(self, planes: torch.Tensor) -> List[torch.Tensor]:
        z_off, _ = self.plan.local_planes(self.rank)
        out = []
        for up in (False, True):
            ov = self.plan.overlap(self.rank, up)
            if ov is not None:
                out.append(planes[ov[0] - z_off: ov[0] - z_off + ov[1]])
        return out

    def _exchange(self, planes: torch.Tensor, box: torch.Tensor) -> None:
        if self.world == 1 or self.skip_exchange:
            return
        if self._fused is not None:
            self._fused.exchange(planes, self.plan, self.c, box)
        elif self._peer is not None:
            self._peer.exchange(planes, self.plan, self.ops.mpm_planes_add)
        else:
            exchange_sum(planes, self.plan, self.rank, self.group, add=self.ops.mpm_planes_add, scratch=self._scratch)

    def _run(self, key: str, body) -> None:
        """eager, or captured once into a CUDA graph (kernels, peer copies and flag kernels) and
        replayed: at 8 ranks a slab's kernels are short enough to be launch-bound."""
        if not self.use_graphs:
            body()
            return
        g = self._graphs.get(key)
        if g is None:
            if key == "bwd":
                keep = [t.clone() for t in self.g[0]]
            body()                                   # warm-up on the current stream: workspaces, scratch
            torch.cuda.synchronize(self.dev)
            if key == "bwd":                         # the warm-up consumed the upstream grads
                for dst, src in zip(self.g[0], keep):
                    dst.copy_(src)
                for t in (self.gWe_part, self.gWp_part):
                    t.zero_()
                torch.cuda.synchronize(self.dev)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                body()
            self._graphs[key] = g
        g.replay()

    # ------------------------------------------------------------------ rollout
    def forward(self):
        self._run("fwd", self._forward_body)
        S = self.S
        return self.x[S], self.v[S], self.C[S], self.F[S]

    def _mpm_fwd(self, c, s: int, grid_op: bool = True) -> None:
        ops = self.ops
        if grid_op:
            ops.mpm_grid_op_box(c, self.box[s], self.grid[s], None)
        ops.mpm_g2p_from(c, self.x[s], self.F[s], self.clip, self.enabled, self.grid[s],
                         (self.x[s + 1], self.v[s + 1], self.C[s + 1], self.Ftrial[s]), self.box[s + 1])

    def _p2g(self, c, s: int) -> None:
        self.ops.mpm_p2g_to(c, self.x[s], self.v[s], self.C[s], self.stress[s], self.vol, self.rho, self.enabled,
                            self.grid[s], box=self.box[s])

    def _elastic_p2g(self, c, s: int, lo: int, hi: int) -> None:
        """elasticity (+ P2G, fused or as a second kernel) of the particle range [lo, hi)"""
        ops = self.ops
        if hi <= lo:
            return
        sl = slice(lo, hi)
        if self.fuse:
            ops.constitutive_fwd_p2g(self.kind_e, self.F[s][sl], self.We_st, self.stress[s][sl], c, self.x[s][sl],
                                     self.v[s][sl], self.C[s][sl], self.vol[sl], self.rho[sl], self.enabled[sl],
                                     self.grid[s], box=self.box[s])
        else:
            ops.constitutive_fwd_staged(self.kind_e, self.F[s][sl], self.We_st, 0.0, self.stress[s][sl])
            ops.mpm_p2g_to(c, self.x[s][sl], self.v[s][sl], self.C[s][sl], self.stress[s][sl], self.vol[sl],
                           self.rho[sl], self.enabled[sl], self.grid[s], box=self.box[s])

    def _forward_body(self) -> None:
        ops, c = self.ops, self.c
        main = torch.cuda.current_stream(self.dev)
        windowed = self._fused is not None
        legacy = self.overlap in ("scatter", "full")

        def clear(s):
            ops.mpm_box_clear(c, self.box[s], self.grid[s], None, self.box[s + 1])
            if self.world > 1 and not windowed:
                for vw in self._overlap_views(self.grid[s]):       # whole shared planes are added: must be valid
                    vw.zero_()
        clear(0)
        for s in range(self.S):
            if legacy:
                ops.constitutive_fwd_staged(self.kind_e, self.F[s], self.We_st, 0.0, self.stress[s])
                self._p2g(self.c_bnd, s)                            # particles touching a shared plane first
                self._side.wait_stream(main)
                with torch.cuda.stream(self._side):
                    self._exchange(self.grid[s], self.box[s])       # the ONE exchange of the substep ...
                self._p2g(self.c_int, s)                            # ... under the interior particles' work
                if self.overlap == "full":
                    self._mpm_fwd(self.c_int, s)
                    main.wait_stream(self._side)
                    self._mpm_fwd(self.c_bnd, s)
                else:
                    main.wait_stream(self._side)
                    self._mpm_fwd(c, s)
            elif self.overlap == "split":
                self._elastic_p2g(c, s, 0, self.split_fwd)          # boundary range: elasticity + P2G
                self._side.wait_stream(main)
                with torch.cuda.stream(self._side):
                    self._exchange(self.grid[s], self.box[s])       # the ONE exchange of the substep ...
                self._elastic_p2g(self.c_chk, s, self.split_fwd, self.n)   # ... under the interior range's kernel
                main.wait_stream(self._side)
                self._mpm_fwd(c, s)
            else:
                self._elastic_p2g(c, s, 0, self.n)
                self._exchange(self.grid[s], self.box[s])           # the ONE exchange of the substep
                self._mpm_fwd(c, s)
            if s + 1 < self.S:        # next substep's grid clear on a side branch next to the plasticity kernel
                self._side.wait_stream(main)
                with torch.cuda.stream(self._side):
                    clear(s + 1)
            ops.constitutive_fwd_staged(ops.KIND_PLASTIC, self.Ftrial[s], self.Wp_st, self.alpha, self.F[s + 1])
            if s + 1 < self.S:
                main.wait_stream(self._side)

    def backward(self, gx, gv, gC, gF):
        """dL/d(local state[S]) -> dL/d(local state[0]); gWe / gWp hold the ALL-REDUCED weight gradients."""
        ops, c = self.ops, self.c
        for dst, src in zip(self.g[0], (gx, gv, gC, gF)):
            dst.zero_() if src is None else dst.copy_(src)
        for t in self.gWe + self.gWp:
            t.zero_()
        self._run("bwd", self._backward_body)
        ops.constitutive_reduce_partials(self.gWe_part, *self.gWe)
        ops.constitutive_reduce_partials(self.gWp_part, *self.gWp)
        if self.world > 1:
            flat = torch.cat([t.reshape(-1) for t in self.gWe + self.gWp])
            dist.all_reduce(flat, group=self.group)
            o = 0
            for t in self.gWe + self.gWp:
                t.copy_(flat[o:o + t.numel()].view_as(t))
                o += t.numel()
        return self.g[self.S % 2]

    def _plastic_bwd_scatter(self, c, s: int, lo: int, hi: int, gx2, gv2, gC2, gF_next, adj=None) -> None:
        """plasticity backward (+ adjoint scatter of G2P, fused or as a second kernel) of [lo, hi)"""
        ops = self.ops
        if hi <= lo:
            return
        adj = self.adj if adj is None else adj
        sl = slice(lo, hi)
        if self.fuse:
            ops.constitutive_bwd_scatter(self.Ftrial[s][sl], self.Wp_st, self.alpha, gF_next[sl], self.gFtrial[sl],
                                         self.gWp_part, c, self.x[s][sl], self.F[s][sl], self.clip[sl],
                                         self.enabled[sl], self.v[s + 1][sl], gx2[sl], gv2[sl], gC2[sl], adj,
                                         box=self.box[s])
        else:
            ops.constitutive_bwd_partials(ops.KIND_PLASTIC, self.Ftrial[s][sl], self.Wp_st, self.alpha, gF_next[sl],
                                          self.gFtrial[sl], self.gWp_part)
            ops.mpm_bwd_scatter(c, self.x[s][sl], self.F[s][sl], self.clip[sl], self.enabled[sl], self.v[s + 1][sl],
                                gx2[sl], gv2[sl], gC2[sl], self.gFtrial[sl], adj, box=self.box[s])

    def _backward_body(self) -> None:
        ops, c = self.ops, self.c
        main = torch.cuda.current_stream(self.dev)
        windowed = self._fused is not None
        legacy = self.overlap in ("scatter", "full")
        cur, nxt = 0, 1

        def clear(s, adj=None):      # the adjoint scatter target, cleared over a substep's box inside the backward
            adj = self.adj if adj is None else adj
            ops.mpm_box_clear(c, self.box[s], adj, None, None)
            if self.world > 1 and not windowed:
                for vw in self._overlap_views(adj):
                    vw.zero_()
        if self.fuse_gather and self.overlap in ("none", "split"):
            # [plasticity bwd + adjoint scatter] [exchange] [node pass] [adjoint gather + elasticity bwd]; two
            # scatter targets in turn, so that the clear for substep s - 1 runs beside substep s (as
            # FusedFrame._sim_backward, fuse_bwd = "full").  split: the exchange (and that clear) run on the side
            # stream under the interior range's plasticity kernel.
            adjs = (self.adj, self.adj2)
            split = self.overlap == "split"
            clear(self.S - 1, adjs[(self.S - 1) & 1])
            for s in reversed(range(self.S)):
                gx2, gv2, gC2, gF_next = self.g[cur]
                gxo, gvo, gCo, gFo = self.g[nxt]
                adj = adjs[s & 1]
                if s > 0:            # the other target's clear: beside the first kernel of this substep
                    self._side.wait_stream(main)
                    with torch.cuda.stream(self._side):
                        clear(s - 1, adjs[(s - 1) & 1])
                if split:
                    self._plastic_bwd_scatter(c, s, 0, self.split_bwd, gx2, gv2, gC2, gF_next, adj)
                    self._side.wait_stream(main)
                    with torch.cuda.stream(self._side):
                        self._exchange(adj, self.box[s])
                    self._plastic_bwd_scatter(self.c_chk, s, self.split_bwd, self.n, gx2, gv2, gC2, gF_next, adj)
                    main.wait_stream(self._side)
                else:
                    self._plastic_bwd_scatter(c, s, 0, self.n, gx2, gv2, gC2, gF_next, adj)
                    self._exchange(adj, self.box[s])
                ops.mpm_bwd_node(c, self.grid[s], adj, self.box[s])
                ops.constitutive_bwd_gather(self.kind_e, self.We_st, self.gWe_part, c, self.x[s], self.v[s], self.C[s],
                                            self.F[s], self.stress[s], self.vol, self.rho, self.clip, self.enabled,
                                            self.v[s + 1], self.C[s + 1], gx2, gv2, gC2, self.gFtrial, self.grid[s],
                                            adj, gxo, gvo, gCo, gFo, self.gstress, box=self.box[s])
                if s > 0 and not split:
                    main.wait_stream(self._side)
                cur, nxt = nxt, cur
            return
        clear(self.S - 1)
        for s in reversed(range(self.S)):
            gx2, gv2, gC2, gF_next = self.g[cur]
            gxo, gvo, gCo, gFo = self.g[nxt]

            def scatter(cc):
                ops.mpm_bwd_scatter(cc, self.x[s], self.F[s], self.clip, self.enabled, self.v[s + 1], gx2, gv2, gC2,
                                    self.gFtrial, self.adj, box=self.box[s])

            def node_gather(cc):
                ops.mpm_bwd_node_gather(cc, self.x[s], self.v[s], self.C[s], self.F[s], self.stress[s], self.vol,
                                        self.rho, self.clip, self.enabled, self.v[s + 1], self.C[s + 1], gx2, gv2,
                                        gC2, self.gFtrial, self.grid[s], self.adj, self.box[s], gxo, gvo, gCo, gFo,
                                        self.gstress)
            if legacy:
                ops.constitutive_bwd_partials(ops.KIND_PLASTIC, self.Ftrial[s], self.Wp_st, self.alpha, gF_next,
                                              self.gFtrial, self.gWp_part)
                scatter(self.c_bnd)
                self._side.wait_stream(main)
                with torch.cuda.stream(self._side):
                    self._exchange(self.adj, self.box[s])           # the ONE exchange of the backward substep
                scatter(self.c_int)
                if self.overlap == "full":
                    node_gather(self.c_int)
                    main.wait_stream(self._side)
                    node_gather(self.c_bnd)
                else:
                    main.wait_stream(self._side)
                    node_gather(c)
            elif self.overlap == "split":
                self._plastic_bwd_scatter(c, s, 0, self.split_bwd, gx2, gv2, gC2, gF_next)
                self._side.wait_stream(main)
                with torch.cuda.stream(self._side):
                    self._exchange(self.adj, self.box[s])           # the ONE exchange of the backward substep
                self._plastic_bwd_scatter(self.c_chk, s, self.split_bwd, self.n, gx2, gv2, gC2, gF_next)
                main.wait_stream(self._side)
                node_gather(c)
            else:
                self._plastic_bwd_scatter(c, s, 0, self.n, gx2, gv2, gC2, gF_next)
                self._exchange(self.adj, self.box[s])
                node_gather(c)
            if s > 0:          # adj was last read by the gather: clear it for substep s - 1 next to the elasticity bwd
                self._side.wait_stream(main)
                with torch.cuda.stream(self._side):
                    clear(s - 1)
            ops.constitutive_bwd_partials(self.kind_e, self.F[s], self.We_st, 0.0, self.gstress, gFo,
                                          self.gWe_part, accumulate_gF=True)
            if s > 0:
                main.wait_stream(self._side)
            cur, nxt = nxt, cur

    # ------------------------------------------------------------------ ownership
    def check_drift(self) -> None:
        """Raises if, during the last forward(), a particle's stencil needed a grid plane this rank
        does not hold (it moved more than the halo allowance k within the frame): its node visits
        were dropped on the device, so state and gradients of that frame are wrong.  The count is
        reduced on the device by G2P next to the node box (box[:, 7]).  Synchronises."""
        lost = int(self.box[:, 7].sum().item()) if self.n else 0
        if lost:
            raise RuntimeError(f"SlabRollout rank {self.rank}: {lost} particle-substeps left the slab's planes "
                               f"(halo_k={self.k}); re-run with a larger halo_k or repartition more often")

    def drift_ok(self) -> bool:
        """True while every local particle still sits inside the planes this rank holds."""
        if self.n == 0:
            return True
        base = SlabPlan.base_plane(self.x[self.S][:, 2], self.G)
        lo, hi = self.plan.tolerated(self.rank)
        return bool(((base >= lo) & (base < hi)).all().item())

    def gather_global(self, t_local: torch.Tensor, n_global: int) -> torch.Tensor:
        """all-gather a per-particle array back into GLOBAL particle order (every rank gets it)."""
        if self.world == 1:
            out = torch.empty((n_global,) + tuple(t_local.shape[1:]), device=t_local.device, dtype=t_local.dtype)
            out[self.ids] = t_local
            return out
        counts = [torch.zeros(1, dtype=torch.int64, device=self.dev) for _ in range(self.world)]
        dist.all_gather(counts, torch.tensor([self.n], dtype=torch.int64, device=self.dev), group=self.group)
        counts = [int(c_.item()) for c_ in counts]
        mx = max(counts)
        pad_ids = torch.zeros(mx, dtype=torch.int64, device=self.dev)
        pad_ids[:self.n] = self.ids
        pad = torch.zeros((mx,) + tuple(t_local.shape[1:]), device=self.dev, dtype=t_local.dtype)
        pad[:self.n] = t_local
        all_ids = [torch.empty_like(pad_ids) for _ in range(self.world)]
        all_t = [torch.empty_like(pad) for _ in range(self.world)]
        dist.all_gather(all_ids, pad_ids, group=self.group)
        dist.all_gather(all_t, pad, group=self.group)
        out = torch.empty((n_global,) + tuple(t_local.shape[1:]), device=self.dev, dtype=t_local.dtype)
        for r in range(self.world):
            out[all_ids[r][:counts[r]]] = all_t[r][:counts[r]]
        return out

    def repartition(self, n_global: int) -> None:
        """Frame boundary: particles change owner (migration) and the slabs are re-balanced.
        state[S] becomes the new state[0].  Raises if the finished frame violated the drift bound."""
        self.check_drift()
        S = self.S
        full = [self.gather_global(t, n_global) for t in
                (self.x[S], self.v[S], self.C[S], self.F[S], self.vol, self.rho, self.clip, self.enabled)]
        self.distribute(*full)
