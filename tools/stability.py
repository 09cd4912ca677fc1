import sys, os, math
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
sys.argv = [sys.argv[0]]
import bench
from neuma_b200 import ops, synthetic as S
from neuma_b200.pipeline import FusedFrame
dev = torch.device("cuda:0")
w = bench.load_weights()
for name in ["jellyduck-100k-128", "bouncyball-50k-64"]:
    spec = S.WORKLOADS[name]
    p = S.make_particles(spec, 42)
    d = {k: v.to(dev) for k, v in p.items()}
    for dt in [1e-3, 5e-4, 2.5e-4, 1e-4]:
        cc = ops.mpm_const(spec["G"], dt, 1, (0.0, -9.8, 0.0), 6e-7, "noslip")
        eng = FusedFrame(spec["n"], cc, (d["vol"], d["rho"], d["clip"], d["enabled"]), [t.to(dev) for t in w["elastic"]],
                         [t.to(dev) for t in w["plastic"]], w["alpha"], 50, dev, use_graphs=False)
        eng.load_state(d["x"], d["v"], d["C"], d["F"])
        rep = []
        for k in range(4):
            x, v, C, F = eng.sim_forward()
            rep.append((bool(torch.isfinite(F).all()), float(v.abs().max()), float((F - torch.eye(3, device=dev)).abs().max())))
            eng.load_state(x.clone(), v.clone(), C.clone(), F.clone())
        print(name, "dt", dt, rep)
