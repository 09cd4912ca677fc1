import os, sys, math
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from neuma_b200 import ops, synthetic as S
from neuma_b200.pipeline import FusedFrame
from neuma_b200.glue import Bindings
dev = torch.device("cuda:0")
spec = dict(n=3000, G=32, box=((0.3, 0.7),) * 3)
p = S.make_particles(spec, 1)
g = S.make_gaussians(p["x"], 1000, 32, 2)
z = np.load(os.path.join(ROOT, "tests", "golden", "material_jelly.npz"))
W = {t: [torch.from_numpy(z[f"param_{t}.{l}.fc.weight"]).float() for l in ("layers.0", "layers.1", "final_layer")] for t in ("elastic", "plastic")}
alpha, Ssub, IM = float(z["alpha"]), 2, 64
cam = S.make_camera(IM, IM)
tan = math.tan(cam.FoVx * 0.5)
settings = ops.raster_settings(IM, IM, tan, tan, [1.0, 1.0, 1.0], 1.0, cam.world_view_transform, cam.full_proj_transform, 3, cam.camera_center)
cc = ops.mpm_const(32, 1e-3, 1, (0.0, -9.8, 0.0), 6e-7, "noslip")
d = {k: v.to(dev) for k, v in p.items()}
F0 = p["F"] + 0.02 * torch.randn(spec["n"], 3, 3, generator=torch.Generator().manual_seed(3))
gt = torch.rand(3, IM, IM, generator=torch.Generator().manual_seed(4))
res = {}
for graphs in (False, True):
    for sort in (False, True):
        eng = FusedFrame(spec["n"], cc, (d["vol"], d["rho"], d["clip"], d["enabled"]), [w.to(dev) for w in W["elastic"]],
                         [w.to(dev) for w in W["plastic"]], alpha, Ssub, dev,
                         bind=Bindings(g["rowptr"].to(dev), g["col"].to(dev), g["w"].to(dev), spec["n"]),
                         gaussians={k: g[k].to(dev) for k in ("cov6", "opac", "shs")}, settings=settings, use_graphs=graphs, sort=sort)
        for rep in range(2):
            eng.load_state(d["x"], d["v"], d["C"], F0.to(dev))
            eng.zero_weight_grads()
            loss, g0 = eng.frame_fwd_bwd(gt.to(dev), d["x"].clone(), g["xyz"].to(dev))
            torch.cuda.synchronize()
            res[(graphs, sort, rep)] = (float(loss), [float(t.norm()) for t in eng.gWe + eng.gWp], [float(t.norm()) for t in g0])
for k, v in res.items():
    print(k, v)
