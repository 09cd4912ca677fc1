"""FusedFrame (neuma_b200/pipeline.py): CUDA-graph replay == eager launches, from the very first
frame, with and without the per-frame cell sort; particle order in == particle order out."""
import math
import os

import numpy as np
import pytest
import torch

from util import rel_err

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _setup(dev, n=3000, G=32, IM=64):
    from neuma_b200 import ops, synthetic as S
    from neuma_b200.glue import Bindings
    spec = dict(n=n, G=G, box=((0.3, 0.7),) * 3)
    p = S.make_particles(spec, 1)
    g = S.make_gaussians(p["x"], n // 3, G, 2)
    z = np.load(os.path.join(ROOT, "tests", "golden", "material_jelly.npz"))
    W = {t: [torch.from_numpy(z[f"param_{t}.{l}.fc.weight"]).float().to(dev) for l in ("layers.0", "layers.1", "final_layer")]
         for t in ("elastic", "plastic")}
    cam = S.make_camera(IM, IM)
    tan = math.tan(cam.FoVx * 0.5)
    settings = ops.raster_settings(IM, IM, tan, tan, [1.0, 1.0, 1.0], 1.0, cam.world_view_transform,
                                   cam.full_proj_transform, 3, cam.camera_center)
    cc = ops.mpm_const(G, 1e-3, 1, (0.0, -9.8, 0.0), 6e-7, "noslip")
    d = {k: v.to(dev) for k, v in p.items()}
    F0 = (p["F"] + 0.02 * torch.randn(n, 3, 3, generator=torch.Generator().manual_seed(3))).to(dev)
    gt = torch.rand(3, IM, IM, generator=torch.Generator().manual_seed(4)).to(dev)
    bind = Bindings(g["rowptr"].to(dev), g["col"].to(dev), g["w"].to(dev), n)
    gs = {k: g[k].to(dev) for k in ("cov6", "opac", "shs")}
    return spec, cc, d, W, float(z["alpha"]), settings, bind, gs, F0, gt, g["xyz"].to(dev)


def test_graph_replay_equals_eager_from_the_first_frame(cuda_dev):
    from neuma_b200.pipeline import FusedFrame
    spec, cc, d, W, alpha, settings, bind, gs, F0, gt, kxyz = _setup(cuda_dev)
    runs = {}
    for graphs in (False, True):
        for sort in (False, True):
            eng = FusedFrame(spec["n"], cc, (d["vol"], d["rho"], d["clip"], d["enabled"]), W["elastic"], W["plastic"],
                             alpha, 2, cuda_dev, bind=bind, gaussians=gs, settings=settings, use_graphs=graphs, sort=sort)
            for frame in range(2):
                eng.load_state(d["x"], d["v"], d["C"], F0)
                eng.zero_weight_grads()
                loss, g0 = eng.frame_fwd_bwd(gt, d["x"].clone(), kxyz)
                runs[(graphs, sort, frame)] = (loss.clone(), [t.clone() for t in g0], [t.clone() for t in eng.gWe + eng.gWp])
    ref = runs[(False, False, 0)]
    for key, (loss, g0, gW) in runs.items():
        assert abs(float(loss) - float(ref[0])) <= 1e-6 * abs(float(ref[0])), key
        for a, b in zip(g0, ref[1]):
            assert rel_err(a, b) < 1e-4, key                     # fp32 atomics: order differs between runs
        for a, b in zip(gW, ref[2]):
            assert rel_err(a, b) < 1e-4, key


def test_set_weights_restages(cuda_dev):
    from neuma_b200.pipeline import FusedFrame
    spec, cc, d, W, alpha, settings, bind, gs, F0, gt, kxyz = _setup(cuda_dev, n=1000)
    eng = FusedFrame(spec["n"], cc, (d["vol"], d["rho"], d["clip"], d["enabled"]), W["elastic"], W["plastic"], alpha, 2,
                     cuda_dev, bind=bind, gaussians=gs, settings=settings)
    eng.load_state(d["x"], d["v"], d["C"], F0)
    a = [t.clone() for t in eng.sim_forward()]
    eng.set_weights(We=[1.5 * w for w in W["elastic"]])
    b = [t.clone() for t in eng.sim_forward()]
    assert rel_err(a[1], b[1]) > 1e-4                            # stiffer material: different velocities
    eng.set_weights(We=W["elastic"])
    c = eng.sim_forward()
    for x, y in zip(a, c):
        assert rel_err(x, y) < 1e-5
