"""Z-slab decomposition (neuma_b200/slab.py): host logic + the summing halo exchange, on CPU
with the gloo backend (world_size 2 and 3).  The per-rank "kernels" here are the CPU oracle's
p2g -- only a checker for the partition / exchange plumbing; the GPU path is covered by
tests/test_slab_gpu.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from neuma_b200.slab import SlabPlan, exchange_sum
from oracle import mpm_oracle as O
from util import make_cloud


def test_plan_balanced_and_thick_enough():
    G = 64
    z = torch.rand(20000) * 0.4 + 0.3                      # compact cloud: most planes are empty
    base = SlabPlan.base_plane(z, G)
    hist = torch.bincount(base, minlength=G).numpy()
    for world in (1, 2, 4, 8):
        plan = SlabPlan.from_histogram(hist, world, k=1)
        assert plan.b[0] == 0 and plan.b[-1] == G and len(plan.b) == world + 1
        thick = np.diff(plan.b)
        assert world == 1 or thick.min() >= SlabPlan.min_thickness(1)
        owner = plan.owner_of(base)
        counts = torch.bincount(owner, minlength=world)
        assert counts.sum() == 20000
        if world in (2, 4):                                # equal-count slabs (quantised to whole planes)
            assert counts.max() - counts.min() <= 2 * hist.max()
        for r in range(world):
            z_off, nz = plan.local_planes(r)
            lo, hi = plan.tolerated(r)
            assert z_off <= max(lo, 0) and hi + 2 <= z_off + nz + (0 if r < world - 1 else 2)
            # neighbours agree on the shared block
            if r + 1 < world:
                assert plan.overlap(r, True) == plan.overlap(r + 1, False)


def test_plan_rejects_bad_input():
    with pytest.raises(ValueError):
        SlabPlan(32, [0, 3, 32], k=1)                      # 3 planes < 2 + 2k
    with pytest.raises(ValueError):
        SlabPlan.from_histogram(np.ones(16), 8, k=1)       # 8 slabs x 4 planes > 16
    assert SlabPlan.base_plane(torch.tensor([0.01]), 32).item() == 0   # int() truncates: -0.18 -> 0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _pack(mv, m):
    """reference layout [x][y][z] -> packed plane-major [z][y][x][4] (include/neuma_b200.h)"""
    return torch.cat([mv, m[..., None]], -1).permute(2, 1, 0, 3).contiguous()


def _worker(rank, world, port, G, n, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        d = make_cloud(n, G, seed=3, lo=0.2, hi=0.8, dtype=torch.float64)
        co = O.MPMConst(num_grids=G, dt=1e-3, bound=1, eps=6e-7, bc="noslip")
        base = SlabPlan.base_plane(d["x"][:, 2], G)
        plan = SlabPlan.from_histogram(torch.bincount(base, minlength=G).numpy(), world, k=1)
        mine = plan.owner_of(base) == rank
        # particles drift by < k = 1 plane before re-partitioning: emulate by moving some of mine
        z_off, nz = plan.local_planes(rank)
        loc = {k: t[mine] for k, t in d.items()}
        mv, m = O.p2g(co, loc["x"], loc["v"], loc["C"], loc["stress"], loc["vol"], loc["rho"], loc["enabled"])
        mine_planes = _pack(mv, m)[z_off:z_off + nz].contiguous()
        exchange_sum(mine_planes, plan, rank)                                   # the ONE exchange
        gmv, gm = O.p2g(co, d["x"], d["v"], d["C"], d["stress"], d["vol"], d["rho"], d["enabled"])
        ref = _pack(gmv, gm)[z_off:z_off + nz]
        # planes this rank's particles can read: [b[r], b[r+1] + 2): must hold the GLOBAL totals
        lo, hi = plan.b[rank] - z_off, min(G, plan.b[rank + 1] + 2) - z_off
        err = float((mine_planes[lo:hi] - ref[lo:hi]).abs().max() / ref.abs().max())
        # weight-gradient style all-reduce across the slabs
        t = torch.tensor([float(mine.sum())], dtype=torch.float64)
        dist.all_reduce(t)
        out_q.put((rank, err, int(t.item())))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_halo_exchange_sums_to_global_grid(world):
    G, n = 32, 6000
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, G, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err, total in res:
        assert err < 1e-12, (rank, err)
        assert total == n


def test_split_point_sits_on_whole_tile_rounds():
    """overlap="split": two launches over [0, split) and [split, n) must not cost more tile rounds of the
    persistent constitutive kernels than one launch over [0, n)."""
    from neuma_b200.slab import split_on_tile_rounds
    sms, tile = 148, 128
    rounds = lambda cnt, pipes: -(-(-(-cnt // tile)) // (sms * pipes))      # ceil(ceil(cnt / tile) / (sms * pipes))
    for pipes in (2, 3):
        for n, nb in ((127284, 39251), (500000, 41000), (20000, 7000), (1000, 0), (56832, 56832), (300000, 299999)):
            s = split_on_tile_rounds(nb, n, sms, pipes, tile)
            assert nb <= s <= n
            assert s == n or s % (sms * pipes * tile) == 0
            assert rounds(s, pipes) + rounds(n - s, pipes) == rounds(n, pipes) or s == n
