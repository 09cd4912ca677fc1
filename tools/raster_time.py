"""Render forward + backward of the bench workload (50 k Gaussians, 800 x 800) a few times: the target of
   ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/raster_launches.csv python tools/raster_time.py"""
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from neuma_b200 import ops, synthetic as S
dev = torch.device("cuda:0")
spec = S.WORKLOADS["jellyduck-100k-128"]
p = S.make_particles(spec, 42)
gs = S.make_gaussians(p["x"], spec["K"], spec["G"], 43)
cam = S.make_camera(800, 800)
tan = math.tan(cam.FoVx * 0.5)
cs = ops.raster_settings(800, 800, tan, tan, [1.0, 1.0, 1.0], 1.0, cam.world_view_transform, cam.full_proj_transform, 3,
                         cam.camera_center)
means, shs, opac, cov = [gs[k].to(dev) for k in ("xyz", "shs", "opac", "cov6")]
G = torch.randn(3, 800, 800, device=dev)
for _ in range(int(os.environ.get("REPS", "4"))):
    img, radii = ops.raster_fwd(cs, means, shs, None, opac, cov)
    ops.raster_bwd(cs, means, shs, None, opac, cov, radii, G)
torch.cuda.synchronize()
keys, ids, ranges = ops.raster_binning(img, 2500)
cnt = (ranges[:, 1] - ranges[:, 0]).cpu()
print("instances", int(cnt.sum()), "tiles>0", int((cnt > 0).sum()), "max/tile", int(cnt.max()), "tiles>2048", int((cnt > 2048).sum()),
      "tiles>4096", int((cnt > 4096).sum()))
