"""A/B of the two P2G kernels at BASELINE's four particle / grid shapes (cell-ordered particles, compact
grid slot, CUDA events, L2 flushed between launches AND warm):
    python tools/p2g_ab.py                 # direct: one red.global.add.v4.f32 per node visit
    NEUMA_P2G=tiled python tools/p2g_ab.py  # shared-memory tile + flush (the north_star's shape)
Also checks the result of the selected kernel against the other one's (both in one process is not
possible: the choice is latched at the first call), via a checksum printed for both."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from neuma_b200 import ops, synthetic as S
dev = torch.device("cuda:0")
mode = os.environ.get("NEUMA_P2G", "direct")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for name in ("bouncyball-50k-64", "jellyduck-100k-128", "claycat-200k-128", "block-1m-256"):
    spec = S.WORKLOADS[name]
    p = S.make_particles(spec, 42)
    g = torch.Generator().manual_seed(0)
    n, G = spec["n"], spec["G"]
    d = {k: t.to(dev) for k, t in p.items()}
    d["C"] = (0.5 * torch.randn(n, 3, 3, generator=g)).to(dev)
    stress = (50 * torch.randn(n, 3, 3, generator=g)).to(dev)
    c0 = ops.mpm_const(G, spec["dt"], 1, (0.0, -9.8, 0.0), 6e-7, "noslip")
    perm = ops.mpm_cell_order(c0, d["x"])
    d = {k: ops.gather_rows(t, perm) for k, t in d.items()}
    stress = ops.gather_rows(stress, perm)
    cap = 1 << 21 if G <= 128 else 1 << 24
    c = ops.mpm_const(G, spec["dt"], 1, (0.0, -9.8, 0.0), 6e-7, "noslip", cap_nodes=cap)
    grid = torch.zeros(cap, 4, device=dev)
    box = torch.zeros(2, 8, dtype=torch.int32, device=dev)
    ops.mpm_box_from_x(c, d["x"], d["enabled"], box[0])

    def run():
        ops.mpm_p2g_to(c, d["x"], d["v"], d["C"], stress, d["vol"], d["rho"], d["enabled"], grid, box=box[0])

    def clear_run():
        ops.mpm_box_clear(c, box[0], grid, None, box[1])
        run()
    res = {}
    for tag, fn, fl in (("p2g_cold", run, True), ("p2g_warm", run, False), ("clear+p2g_cold", clear_run, True),
                        ("clear+p2g_warm", clear_run, False)):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(20):
            if fl:
                flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); b.synchronize()
            ts.append(a.elapsed_time(b) * 1e3)
        res[tag] = sum(ts) / len(ts)
    ops.mpm_box_clear(c, box[0], grid, None, box[1])
    run()
    torch.cuda.synchronize()
    chk = grid.double().sum(0).tolist()
    dense = 108 * n + 44 * G ** 3
    print(f"{mode:7s} {name:20s} " + "  ".join(f"{k}={v:7.2f}us" for k, v in res.items()) +
          f"  dense-equivalent {dense / res['clear+p2g_cold'] / 1e3:7.1f} GB/s (cold) {dense / res['clear+p2g_warm'] / 1e3:7.1f} (warm)"
          f"  checksum(mv,m)=({chk[0]:.6e},{chk[1]:.6e},{chk[2]:.6e},{chk[3]:.9e})")
