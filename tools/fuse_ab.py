"""A/B of the G2P / adjoint-gather fusions at the headline shape (jellyduck 100 k / 128^3, cell-ordered
particles, compact slots, weights of the golden jelly material): each stage as its own kernel against the
fused kernel, CUDA events over graph replays of 20 back-to-back launches on one substep's buffers, L2 warm (inside a frame
the producer kernel has just written the operands).
    python tools/fuse_ab.py [workload]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from neuma_b200 import ops, synthetic as S
from neuma_b200.pipeline import FusedFrame

dev = torch.device("cuda:0")
name = sys.argv[1] if len(sys.argv) > 1 else "jellyduck-100k-128"
spec = S.WORKLOADS[name]
n, G = spec["n"], spec["G"]
p = S.make_particles(spec, 42)
d = {k: t.to(dev) for k, t in p.items()}
z = np.load(os.path.join(ROOT, "tests", "golden", "material_jelly.npz"))
W = {t: [torch.from_numpy(z[f"param_{t}.{l}.fc.weight"]).float().to(dev) for l in ("layers.0", "layers.1", "final_layer")]
     for t in ("elastic", "plastic")}
alpha = float(z["alpha"])
cc = ops.mpm_const(G, spec["dt"], 1, (0.0, -9.8, 0.0), 6e-7, "noslip")
eng = FusedFrame(n, cc, (d["vol"], d["rho"], d["clip"], d["enabled"]), W["elastic"], W["plastic"], alpha, 12, dev,
                 use_graphs=False, fuse="full")
eng.load_state(d["x"], d["v"], d["C"], d["F"])
eng.sim_forward()
for t, g in zip(eng.g[0], (3, 3, 9, 9)):
    t.normal_()
eng._sim_backward(0)
torch.cuda.synchronize()
s = 10
c = eng.c
e = eng
tmp = [torch.empty_like(t) for t in (e.x[0], e.v[0], e.C[0], e.F[0])]
Fo = torch.empty_like(e.F[0])
boxn = torch.zeros(8, dtype=torch.int32, device=dev)
gx2, gv2, gC2, gFn = [torch.randn_like(t) for t in (e.x[0], e.v[0], e.C[0], e.F[0])]
go = [torch.empty_like(t) for t in (e.x[0], e.v[0], e.C[0], e.F[0], e.F[0])]
adj = e.adj


def timeit(fn, reps=20):
    """20 launches captured into ONE CUDA graph (the Python / ctypes launch path costs more than these
    kernels run), replayed 5 times between events."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    side = torch.cuda.Stream(dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        fn()
    torch.cuda.current_stream(dev).wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        g.replay()
    b.record()
    b.synchronize()
    return a.elapsed_time(b) * 1e3 / (5 * reps)


def g2p():
    ops.mpm_g2p_from(c, e.x[s], e.F[s], e.clip, e.enabled, e.grid[s], tmp, boxn, box=e.box[s])


def plast():
    ops.constitutive_fwd_staged(ops.KIND_PLASTIC, e.Ftrial[s], e.Wp_st, alpha, Fo)


def g2p_plast():
    ops.constitutive_fwd_g2p(e.Wp_st, alpha, Fo, c, e.x[s], e.F[s], e.clip, e.enabled, e.grid[s], tmp, box_next=boxn,
                             box=e.box[s])


def gather():
    ops.mpm_bwd_node_gather(c, e.x[s], e.v[s], e.C[s], e.F[s], e.stress[s], e.vol, e.rho, e.clip, e.enabled, e.v[s + 1],
                            e.C[s + 1], gx2, gv2, gC2, gFn, e.grid[s], adj, e.box[s], *go)


def node():
    ops.mpm_bwd_node(c, e.grid[s], adj, e.box[s])


def ebwd():
    ops.constitutive_bwd_partials(e.kind_e, e.F[s], e.We_st, 0.0, go[4], go[3], e.gWe_part, accumulate_gF=True)


def gather_ebwd():
    ops.constitutive_bwd_gather(e.kind_e, e.We_st, e.gWe_part, c, e.x[s], e.v[s], e.C[s], e.F[s], e.stress[s], e.vol,
                                e.rho, e.clip, e.enabled, e.v[s + 1], e.C[s + 1], gx2, gv2, gC2, gFn, e.grid[s], adj,
                                go[0], go[1], go[2], go[3], go[4], box=e.box[s])


res = {}
for k, fn in (("g2p", g2p), ("plasticity", plast), ("g2p;plasticity", lambda: (g2p(), plast())),
              ("g2p+plasticity fused", g2p_plast), ("node", node), ("node;gather", gather), ("elasticity_bwd", ebwd),
              ("node;gather;elasticity_bwd", lambda: (gather(), ebwd())),
              ("node;gather+elasticity_bwd fused", lambda: (node(), gather_ebwd()))):
    res[k] = timeit(fn)
print(name, " ".join(f"[{k}] {v:.1f} us " for k, v in res.items()))
