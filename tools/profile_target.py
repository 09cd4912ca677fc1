"""Runs each hot kernel a few times at the bench workload's shapes -- the target for `ncu --set full`."""
import sys, os, math
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
sys.argv = [sys.argv[0]]
import bench
from neuma_b200 import ops, synthetic as S
from neuma_b200.pipeline import FusedFrame
from neuma_b200.glue import Bindings

dev = torch.device("cuda:0")
wl = bench.build_workload(2)
spec, p, g = wl["spec"], wl["p"], wl["g"]
n, G = spec["n"], spec["G"]
cc = ops.mpm_const(G, spec["dt"], 1, (0.0, -9.8, 0.0), 6e-7, "noslip")
d = {k: v.to(dev) for k, v in p.items()}
cam = wl["cam"]
settings = ops.raster_settings(800, 800, math.tan(cam.FoVx * 0.5), math.tan(cam.FoVy * 0.5), [1.0, 1.0, 1.0], 1.0,
                               cam.world_view_transform, cam.full_proj_transform, 3, cam.camera_center)
eng = FusedFrame(n, cc, (d["vol"], d["rho"], d["clip"], d["enabled"]), [w.to(dev) for w in wl["w"]["elastic"]],
                 [w.to(dev) for w in wl["w"]["plastic"]], wl["w"]["alpha"], 2, dev,
                 bind=Bindings(g["rowptr"].to(dev), g["col"].to(dev), g["w"].to(dev), n),
                 gaussians={k: g[k].to(dev) for k in ("cov6", "opac", "shs")}, settings=settings, use_graphs=False)
eng.load_state(d["x"], d["v"], d["C"], d["F"] + 0.02 * torch.randn_like(d["F"]))
gt = wl["gt"].to(dev)
for _ in range(3):
    eng.zero_weight_grads()
    eng.frame_fwd_bwd(gt, d["x"].clone(), g["xyz"].to(dev))
torch.cuda.synchronize()
print("done")
