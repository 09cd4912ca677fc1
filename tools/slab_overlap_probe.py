"""Does the halo exchange really run under the interior range's kernel?  Eager launches (no graphs),
CUDA events at the fork, after the exchange (side stream) and after the interior kernel (main):
    torchrun --nproc-per-node 2 tools/slab_overlap_probe.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
import bench
from neuma_b200 import synthetic as S
from neuma_b200.slab import SlabRollout
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
spec = S.WORKLOADS["block-1m-256"]
p = S.make_particles(spec, 42)
w = bench.load_weights()
sim = SlabRollout(spec["G"], spec["dt"], w["elastic"], w["plastic"], w["alpha"], 20, dev, halo="fused", overlap="split",
                  use_graphs=False)
sim.distribute(p["x"], p["v"], p["C"], p["F"], p["vol"], p["rho"], p["clip"], p["enabled"])
for _ in range(3):
    sim.forward()
torch.cuda.synchronize(); dist.barrier()
ev = []
orig_ex, orig_ep = sim._exchange, sim._elastic_p2g
def ex(planes, box):
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record(); orig_ex(planes, box); b.record(); ev.append(("ex", a, b))
def ep(c, s, lo, hi):
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record(); orig_ep(c, s, lo, hi); b.record(); ev.append(("ep_bnd" if lo == 0 else "ep_int", a, b))
orig_pb = sim._plastic_bwd_scatter
def pb(c, s, lo, hi, *a_):
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record(); orig_pb(c, s, lo, hi, *a_); b.record(); ev.append(("pb_bnd" if lo == 0 else "pb_int", a, b))
sim._exchange, sim._elastic_p2g, sim._plastic_bwd_scatter = ex, ep, pb
sim.forward()
nf = len(ev)
g = [torch.randn(sim.n, 3, device=dev), None, None, torch.randn(sim.n, 3, 3, device=dev)]
sim.backward(*g)
torch.cuda.synchronize()
who = min(3, world - 1)
for part, lo_, hi_ in (("fwd", 0, nf), ("bwd", nf, len(ev))):
    t0 = ev[lo_][1]
    rows = [(k, t0.elapsed_time(a) * 1e3, t0.elapsed_time(b) * 1e3) for k, a, b in ev[lo_:hi_]]
    if rank == who:
        print(f"rank {rank} {part}: nb={sim.nb} n={sim.n}")
        for k, a, b in rows[6:18]:
            print(f"  {k:7s} start {a:9.1f} us  end {b:9.1f} us  dur {b - a:7.1f}")
        print(f"  whole {part} (eager, S=20): {rows[-1][2] / 1e3:.2f} ms")
dist.barrier(); dist.destroy_process_group()
