"""Developer timing of the individual entry points (CUDA events, L2 flushed between iterations)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from neuma_b200 import ops, build
from util import make_cloud

build.build()
dev = torch.device("cuda:0")
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def timeit(fn, iters=20, warm=3, flush_l2=True):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        if flush_l2:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


for n, G in [(100_000, 128), (200_000, 128), (1_000_000, 256), (50_000, 64)]:
    d = {k: t.to(dev) for k, t in make_cloud(n, G, seed=42).items()}
    cc = ops.mpm_const(G, 1e-3, 1, (0, -9.8, 0), 6e-7, "noslip")
    out = tuple(torch.empty_like(d[k]) for k in ("x", "v", "C", "F"))
    gouts = [torch.randn_like(o) for o in out]
    W0 = torch.randn(64, 13, device=dev) * 0.3
    W1 = torch.randn(64, 64, device=dev) * 0.1
    W2 = torch.randn(9, 64, device=dev) * 0.1
    Ng = G ** 3
    res = {"n": n, "G": G}
    res["clear_us"] = timeit(lambda: ops.mpm_grid_clear(cc, d["x"]))
    res["p2g_us"] = timeit(lambda: ops.mpm_p2g(cc, d["x"], d["v"], d["C"], d["stress"], d["vol"], d["rho"], d["enabled"]))
    res["grid_op_us"] = timeit(lambda: ops.mpm_grid_op(cc, d["x"]))
    res["g2p_lazy_us"] = timeit(lambda: ops.mpm_g2p(cc, d["x"], d["F"], d["clip"], d["enabled"], out=out, lazy_grid_op=True))
    res["g2p_dense_us"] = timeit(lambda: ops.mpm_g2p(cc, d["x"], d["F"], d["clip"], d["enabled"], out=out, lazy_grid_op=False))
    f = lambda: ops.mpm_substep_fwd(cc, d["x"], d["v"], d["C"], d["F"], d["stress"], d["vol"], d["rho"], d["clip"], d["enabled"], out=out)
    res["substep_fwd_us"] = timeit(f)
    res["substep_fwd_hot_us"] = timeit(f, flush_l2=False)
    b = lambda: ops.mpm_substep_bwd(cc, d["x"], d["v"], d["C"], d["F"], d["stress"], d["vol"], d["rho"], d["clip"], d["enabled"], out[1], *gouts)
    res["substep_bwd_us"] = timeit(b)
    res["substep_bwd_hot_us"] = timeit(b, flush_l2=False)
    # ---- rollout path (per-substep grid, box-sparse node passes), caller order vs cell order
    for tag, dd in (("", d), ("_sorted", None)):
        if dd is None:
            perm = ops.mpm_cell_order(cc, d["x"])
            dd = {k: ops.gather_rows(t.contiguous(), perm) for k, t in d.items()}
        grid = torch.zeros(Ng, 4, device=dev); adj = torch.zeros(Ng, 4, device=dev)
        box = torch.zeros(2, 8, dtype=torch.int32, device=dev)
        ops.mpm_box_from_x(cc, dd["x"], dd["enabled"], box[0])
        res["r_clear_us" + tag] = timeit(lambda: ops.mpm_box_clear(cc, box[0], grid, None, box[1]))
        res["r_p2g_us" + tag] = timeit(lambda: ops.mpm_p2g_to(cc, dd["x"], dd["v"], dd["C"], dd["stress"], dd["vol"], dd["rho"], dd["enabled"], grid, box=box[0]))
        res["r_gridop_us" + tag] = timeit(lambda: ops.mpm_grid_op_box(cc, box[0], grid, adj))
        res["r_g2p_us" + tag] = timeit(lambda: ops.mpm_g2p_from(cc, dd["x"], dd["F"], dd["clip"], dd["enabled"], grid, out, box[1], box=box[0]))
        rf = lambda: ops.mpm_rollout_fwd(cc, dd["x"], dd["v"], dd["C"], dd["F"], dd["stress"], dd["vol"], dd["rho"], dd["clip"], dd["enabled"], out, grid, adj, box[0], box[1])
        res["r_fwd_us" + tag] = timeit(rf)
        g5 = [torch.empty_like(dd[k]) for k in ("x", "v", "C", "F", "stress")]
        rb = lambda: ops.mpm_rollout_bwd(cc, dd["x"], dd["v"], dd["C"], dd["F"], dd["stress"], dd["vol"], dd["rho"], dd["clip"], dd["enabled"], out[1], out[2], *gouts, grid, adj, box[0], *g5)
        res["r_bwd_us" + tag] = timeit(rb)
        res["r_p2g_algo_GBs" + tag] = (108 * n + 44 * Ng) / (res["r_clear_us" + tag] + res["r_p2g_us" + tag]) / 1e3
    res["fwd_algo_GBs"] = (260 * n + 84 * Ng) / res["substep_fwd_us"] / 1e3
    res["p2g_algo_GBs"] = (108 * n + 44 * Ng) / (res["clear_us"] + res["p2g_us"]) / 1e3
    res["const_fwd_us"] = timeit(lambda: ops.constitutive_fwd(0, d["F"], W0, W1, W2))
    go = torch.randn_like(d["F"])
    res["const_bwd_us"] = timeit(lambda: ops.constitutive_bwd(0, d["F"], W0, W1, W2, 0.0, go))
    res["const_bwd_nowgrad_us"] = timeit(lambda: ops.constitutive_bwd(0, d["F"], W0, W1, W2, 0.0, go, want_wgrad=False))
    print(json.dumps({k: (round(v, 2) if isinstance(v, float) else v) for k, v in res.items()}))
    sys.stdout.flush()
