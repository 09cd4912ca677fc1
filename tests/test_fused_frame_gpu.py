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


@pytest.mark.parametrize("compact", [True, False])
def test_fuse_modes_agree(cuda_dev, compact):
    """fuse = "full" (3 kernels per substep each way: G2P inside the plasticity kernel, the adjoint gather
    inside the elasticity backward, the next slot prepared by grid_op) and "epilogue" against the un-fused
    launch sequence: same per-particle code, so forward state and loss agree to fp32 atomics order."""
    from neuma_b200.pipeline import FusedFrame
    spec, cc, d, W, alpha, settings, bind, gs, F0, gt, kxyz = _setup(cuda_dev, n=4000)
    runs = {}
    for fuse in (False, "epilogue", "full", "backward-full", True):
        eng = FusedFrame(spec["n"], cc, (d["vol"], d["rho"], d["clip"], d["enabled"]), W["elastic"], W["plastic"],
                         alpha, 4, cuda_dev, bind=bind, gaussians=gs, settings=settings, use_graphs=False, fuse=fuse,
                         compact=compact)
        eng.load_state(d["x"], d["v"], d["C"], F0)
        eng.zero_weight_grads()
        loss, g0 = eng.frame_fwd_bwd(gt, d["x"].clone(), kxyz)
        eng.check()
        runs[fuse] = (loss.clone(), [t.clone() for t in eng.final_state()], [t.clone() for t in g0],
                      [t.clone() for t in eng.gWe + eng.gWp], eng.box.clone())
    ref = runs[False]
    for fuse in ("epilogue", "full", "backward-full", True):
        loss, st, g0, gW, box = runs[fuse]
        assert torch.equal(box[:, :6], ref[4][:, :6]), fuse            # node boxes of every substep
        assert abs(float(loss) - float(ref[0])) <= 1e-6 * abs(float(ref[0])), fuse
        for a, b in zip(st, ref[1]):
            assert rel_err(a, b) < 5e-5, fuse
        for a, b in zip(g0, ref[2]):
            assert rel_err(a, b) < 1e-4, fuse
        for a, b in zip(gW, ref[3]):
            assert rel_err(a, b) < 1e-4, fuse


def test_saved_singular_vectors_equal_recomputed_svd(cuda_dev):
    """save_svd: the backward kernels read the forward's right singular vectors and run only the
    Gram-Schmidt tail of the SVD -- same U, sigma, V, so the same gradients as the full recompute."""
    from neuma_b200.pipeline import FusedFrame
    spec, cc, d, W, alpha, settings, bind, gs, F0, gt, kxyz = _setup(cuda_dev, n=4000)
    runs = {}
    for save in (False, True):
        for fuse in (True, False):
            eng = FusedFrame(spec["n"], cc, (d["vol"], d["rho"], d["clip"], d["enabled"]), W["elastic"], W["plastic"],
                             alpha, 4, cuda_dev, bind=bind, gaussians=gs, settings=settings, use_graphs=False,
                             save_svd=save, fuse=fuse)
            eng.load_state(d["x"], d["v"], d["C"], F0)
            eng.zero_weight_grads()
            loss, g0 = eng.frame_fwd_bwd(gt, d["x"].clone(), kxyz)
            runs[(save, fuse)] = (loss.clone(), [t.clone() for t in g0], [t.clone() for t in eng.gWe + eng.gWp])
    ref = runs[(False, False)]
    for key, (loss, g0, gW) in runs.items():
        assert float(loss) == float(ref[0]), key                # the forward is untouched
        for a, b in zip(g0, ref[1]):
            assert rel_err(a, b) < 1e-4, key
        for a, b in zip(gW, ref[2]):
            assert rel_err(a, b) < 1e-4, key


def test_full_fusion_flags_a_particle_crossing_more_than_a_cell(cuda_dev):
    """fuse="full" clears the next grid slot over the current box dilated by ONE node; a particle that
    leaves that region in one substep must raise check() instead of scattering into stale nodes."""
    from neuma_b200._lib import NeumaError
    from neuma_b200.pipeline import FusedFrame
    spec, cc, d, W, alpha, settings, bind, gs, F0, gt, kxyz = _setup(cuda_dev, n=1000)
    v = d["v"].clone()
    v[:, 0] = 3.0 / (spec["G"] * 1e-3)                # three cells per substep along +x
    eng = FusedFrame(spec["n"], cc, (d["vol"], d["rho"], d["clip"], d["enabled"]), W["elastic"], W["plastic"], alpha, 2,
                     cuda_dev, bind=bind, gaussians=gs, settings=settings, use_graphs=False, grid_margin=12, fuse="full")
    eng.load_state(d["x"], v, d["C"], F0)
    eng.sim_forward()
    with pytest.raises(NeumaError):
        eng.check()
    ok = FusedFrame(spec["n"], cc, (d["vol"], d["rho"], d["clip"], d["enabled"]), W["elastic"], W["plastic"], alpha, 2,
                    cuda_dev, bind=bind, gaussians=gs, settings=settings, use_graphs=False, grid_margin=12,
                    fuse="epilogue")
    ok.load_state(d["x"], v, d["C"], F0)
    ok.sim_forward()
    ok.check()


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
        assert rel_err(x, y) < 5e-5          # two runs of fp32 atomics (different order), per-element metric


def test_two_chained_frames_against_oracle(cuda_dev):
    """finetune.py:360-414 over two rendered frames (S = 2 each) with one FusedFrame per frame:
    loss, dL/d(state0) and dL/dW against torch autograd over the CPU oracle rollout (fp64)."""
    from neuma_b200 import ops, synthetic as S
    from neuma_b200.pipeline import FusedFrame
    from oracle import mpm_oracle as MO, pipeline_oracle as PO, raster_oracle as RO
    n, G, IM, Ssub = 1500, 32, 48, 2
    spec, cc, d, W, alpha, settings, bind, gs, F0, gt, kxyz = _setup(cuda_dev, n=n, G=G, IM=IM)
    center, size = torch.tensor([0.1, 0.0, 0.1]), torch.tensor([0.8, 0.9, 0.8])
    gts = [gt, gt.flip(2)]
    engs = [FusedFrame(n, cc, (d["vol"], d["rho"], d["clip"], d["enabled"]), W["elastic"], W["plastic"], alpha, Ssub,
                       cuda_dev, bind=bind, gaussians=gs, settings=settings, center=center, size=size) for _ in range(2)]
    # ---- forward through both frames, then backward in reverse order
    state = (d["x"], d["v"], d["C"], F0)
    de_prev = ((d["x"] - center.to(cuda_dev)) / size.to(cuda_dev)).clone()
    k_prev = kxyz
    losses = []
    for e, g_ in zip(engs, gts):
        e.load_state(*state)
        e.zero_weight_grads()
        loss, means3D = e.frame_forward(g_, de_prev, k_prev)
        losses.append(loss)
        state = e.final_state()
        de_prev = ((state[0] - center.to(cuda_dev)) / size.to(cuda_dev)).clone()
        k_prev = means3D.clone()
    g1 = engs[1].frame_backward(None)
    g0 = engs[0].frame_backward(g1)
    total = float(losses[0] + losses[1])
    gW = [a + b for a, b in zip(engs[0].gWe + engs[0].gWp, engs[1].gWe + engs[1].gWp)]
    # ---- oracle
    p = {k: v.cpu() for k, v in d.items()}
    cam_s = RO.Settings(IM, IM, settings.tanfovx, settings.tanfovy, np.ones(3, np.float32),
                        np.array(list(settings.viewmatrix), dtype=np.float32).reshape(4, 4),
                        np.array(list(settings.projmatrix), dtype=np.float32).reshape(4, 4), 3,
                        np.array(list(settings.campos), dtype=np.float32))
    B = torch.sparse_csr_tensor(bind.rowptr.cpu().long(), bind.col.cpu().long(), bind.w.cpu().double(),
                                (bind.n_kernels, n)).to_sparse_coo()
    x0 = p["x"].double().requires_grad_(True)
    v0 = p["v"].double().requires_grad_(True)
    C0 = p["C"].double().requires_grad_(True)
    F0d = F0.cpu().double().requires_grad_(True)
    We = [w.cpu().double().requires_grad_(True) for w in W["elastic"]]
    Wp = [w.cpu().double().requires_grad_(True) for w in W["plastic"]]
    co = MO.MPMConst(num_grids=G, dt=1e-3, bound=1, eps=6e-7, bc="noslip")
    # rollout_loss detaches the first de_x_prev from x0, exactly like the engine's caller above
    oloss = PO.rollout_loss(co, x0, v0, C0, F0d, p["vol"], p["rho"], p["clip"], p["enabled"], We, None, Wp, None, alpha,
                            B, kxyz.cpu(), gs["cov6"].cpu(), gs["opac"].cpu(), gs["shs"].cpu(), cam_s,
                            [g_.cpu() for g_ in gts], Ssub, center, size)
    og = torch.autograd.grad(oloss, [x0, v0, C0, F0d] + We + Wp)
    assert abs(total - float(oloss.detach())) / float(oloss.detach()) < 1e-4
    rel = lambda a, b: float((a.cpu().double() - b).norm() / (b.norm() + 1e-30))
    errs = {name: rel(a, b) for a, b, name in zip(g0, og[:4], ["gx", "gv", "gC", "gF"])}
    errs.update({f"gW{i}": rel(a, b) for i, (a, b) in enumerate(zip(gW, og[4:]))})
    print("relative L2 errors vs the fp64 oracle:", {k: f"{v:.2e}" for k, v in errs.items()})
    # fp32 device rollout (atomics, rasterizer alpha / transmittance threshold flips on a 48x48 image)
    # against an fp64 oracle: same bound as the drop-in loop test (test_dropin_gpu.py) for the weights
    assert max(v for k, v in errs.items() if k.startswith("gW")) < 2e-2, errs
    assert max(v for k, v in errs.items() if not k.startswith("gW")) < 5e-2, errs
