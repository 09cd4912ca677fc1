"""GPU parity AT THE BENCHMARKED SHAPES (BASELINE.json configs 3, 4, 5): the CUDA path
through the C ABI against the CPU oracle on the very workloads bench.py times.

  * one substep forward + backward (elasticity -> P2G / grid_op / G2P -> plasticity) at
    100 k / 128^3 (jelly weights) and 200 k / 128^3 (plasticine weights), compact grid slots;
  * one substep forward at 1 M / 256^3;
  * rasterizer forward + backward at 800 x 800 with the 50 k Gaussians of config 3.

Tolerances: forward 1e-4 (north_star), gradients 1e-3, in the per-element metric of
tests/util.py::rel_err.  The measured errors are appended to gpurun_out/parity_errors.jsonl."""
import math
import os

import numpy as np
import pytest
import torch

from oracle import constitutive_oracle as CO, mpm_oracle as MO, raster_oracle as RO
from util import rel_err, record

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _weights(name):
    z = np.load(os.path.join(ROOT, "tests", "golden", f"material_{name}.npz"))
    W = {t: [torch.from_numpy(z[f"param_{t}.{l}.fc.weight"]).float() for l in ("layers.0", "layers.1", "final_layer")]
         for t in ("elastic", "plastic")}
    return W, float(z["alpha"])


def _state(workload, seed=0):
    """the workload's particle cloud with the state of a rollout in flight: F off identity,
    non-zero affine field, velocity noise"""
    from neuma_b200 import synthetic as S
    spec = S.WORKLOADS[workload]
    p = S.make_particles(spec, 42)
    g = torch.Generator().manual_seed(seed)
    n = spec["n"]
    p["F"] = p["F"] + 0.03 * torch.randn(n, 3, 3, generator=g)
    p["C"] = 0.5 * torch.randn(n, 3, 3, generator=g)
    p["v"] = p["v"] + 0.2 * torch.randn(n, 3, generator=g)
    return spec, p


def _gpu_substep(dev, spec, p, W, alpha, gouts=None):
    """elasticity -> MPM substep (compact saved grid) -> plasticity; optionally the reverse pass"""
    from neuma_b200 import ops
    n, G = spec["n"], spec["G"]
    d = {k: t.to(dev) for k, t in p.items()}
    perm = ops.mpm_cell_order(ops.mpm_const(G, spec["dt"], 1, (0.0, -9.8, 0.0), 6e-7, "noslip"), d["x"])
    d = {k: ops.gather_rows(t, perm) for k, t in d.items()}                 # the engine's particle order
    cap = 1 << 21 if G <= 128 else 1 << 24
    c = ops.mpm_const(G, spec["dt"], 1, (0.0, -9.8, 0.0), 6e-7, "noslip", cap_nodes=cap)
    grid = torch.full((cap, 4), float("nan"), device=dev)
    adj = torch.full((cap, 4), float("nan"), device=dev)
    box = torch.zeros(2, 8, dtype=torch.int32, device=dev)
    We = ops.constitutive_stage(*[w.to(dev) for w in W["elastic"]])
    Wp = ops.constitutive_stage(*[w.to(dev) for w in W["plastic"]])
    stress = torch.empty_like(d["F"])
    ops.constitutive_fwd_staged(ops.KIND_ELASTIC, d["F"], We, 0.0, stress)
    ops.mpm_box_from_x(c, d["x"], d["enabled"], box[0])
    out = tuple(torch.empty_like(d[k]) for k in ("x", "v", "C", "F"))
    ops.mpm_rollout_fwd(c, d["x"], d["v"], d["C"], d["F"], stress, d["vol"], d["rho"], d["clip"], d["enabled"], out,
                        grid, None, box[0], box[1])
    F3 = torch.empty_like(d["F"])
    ops.constitutive_fwd_staged(ops.KIND_PLASTIC, out[3], Wp, alpha, F3)
    assert int(box[0, 6]) == 0                                              # the box fitted the compact slot
    inv = ops.invert_perm(perm)
    back = lambda t: ops.gather_rows(t.contiguous(), inv).cpu()
    res = dict(stress=back(stress), x=back(out[0]), v=back(out[1]), C=back(out[2]), F=back(F3))
    if gouts is None:
        return res
    go = [ops.gather_rows(t.to(dev), perm) for t in gouts]                   # dL/d(x', v', C', F'')
    gFt = torch.empty_like(d["F"])
    partp = ops.constitutive_partials(d["F"])
    parte = ops.constitutive_partials(d["F"])
    ops.constitutive_bwd_partials(ops.KIND_PLASTIC, out[3], Wp, alpha, go[3], gFt, partp)
    grads = [torch.empty_like(d[k]) for k in ("x", "v", "C", "F", "F")]
    ops.mpm_box_clear(c, box[0], adj, None, None)
    ops.mpm_rollout_bwd(c, d["x"], d["v"], d["C"], d["F"], stress, d["vol"], d["rho"], d["clip"], d["enabled"],
                        out[1], out[2], go[0], go[1], go[2], gFt, grid, adj, box[0], *grads)
    ops.constitutive_bwd_partials(ops.KIND_ELASTIC, d["F"], We, 0.0, grads[4], grads[3], parte, accumulate_gF=True)
    gWe = [torch.zeros_like(w, device=dev) for w in W["elastic"]]
    gWp = [torch.zeros_like(w, device=dev) for w in W["plastic"]]
    ops.constitutive_reduce_partials(parte, *gWe)
    ops.constitutive_reduce_partials(partp, *gWp)
    res.update(gx=back(grads[0]), gv=back(grads[1]), gC=back(grads[2]), gF=back(grads[3]),
               gWe=[t.cpu() for t in gWe], gWp=[t.cpu() for t in gWp])
    return res


def _oracle_substep(spec, p, W, alpha, gouts=None, dtype=torch.float64):
    co = MO.MPMConst(num_grids=spec["G"], dt=spec["dt"], bound=1, eps=6e-7, bc="noslip")
    cast = lambda t: t.to(dtype)
    rg = gouts is not None
    x, v, C, F = [cast(p[k]).requires_grad_(rg) for k in ("x", "v", "C", "F")]
    We = [cast(w).requires_grad_(rg) for w in W["elastic"]]
    Wp = [cast(w).requires_grad_(rg) for w in W["plastic"]]
    stress = CO.elasticity(F, We)
    x2, v2, C2, F2 = MO.substep_fwd(co, x, v, C, F, stress, cast(p["vol"]), cast(p["rho"]), cast(p["clip"]), p["enabled"])
    F3 = CO.plasticity(F2, Wp, alpha)
    res = dict(stress=stress.detach(), x=x2.detach(), v=v2.detach(), C=C2.detach(), F=F3.detach())
    if rg:
        loss = sum((a * cast(b)).sum() for a, b in zip((x2, v2, C2, F3), gouts))
        g = torch.autograd.grad(loss, [x, v, C, F] + We + Wp)
        res.update(gx=g[0], gv=g[1], gC=g[2], gF=g[3], gWe=list(g[4:7]), gWp=list(g[7:10]))
    return res


@pytest.mark.parametrize("workload,material", [("jellyduck-100k-128", "jelly"), ("claycat-200k-128", "plasticine")])
def test_substep_fwd_bwd_at_bench_shape(cuda_dev, workload, material):
    W, alpha = _weights(material)
    spec, p = _state(workload)
    n = spec["n"]
    g = torch.Generator().manual_seed(3)
    gouts = [torch.randn(n, 3, generator=g), torch.randn(n, 3, generator=g),
             torch.randn(n, 3, 3, generator=g), torch.randn(n, 3, 3, generator=g)]
    got = _gpu_substep(cuda_dev, spec, p, W, alpha, gouts)
    ref = _oracle_substep(spec, p, W, alpha, gouts)
    errs = {}
    for k in ("stress", "x", "v", "C", "F"):
        errs[k] = rel_err(got[k], ref[k])
    for k in ("gx", "gv", "gC", "gF"):
        errs[k] = rel_err(got[k], ref[k])
    for t in ("gWe", "gWp"):
        for i in range(3):
            errs[f"{t}{i}"] = rel_err(got[t][i], ref[t][i])
    record(f"substep_fwd_bwd[{workload}]", **errs)
    for k in ("stress", "x", "v", "C", "F"):
        assert errs[k] < 1e-4, (k, errs)                           # north_star: 1e-4 rel fp32
    for k in ("gx", "gv", "gC", "gF"):
        # fp32 atomics, bf16x2 split GEMMs vs fp64 autograd; torch-CPU fp32 autograd of the oracle is
        # itself at 3.4e-4 (gF) against fp64 in this metric
        assert errs[k] < 1e-3, (k, errs)
    for k in errs:
        if k.startswith("gW"):
            assert errs[k] < 1e-3, (k, errs)


def test_substep_fwd_1m_256(cuda_dev):
    W, alpha = _weights("jelly")
    spec, p = _state("block-1m-256")
    got = _gpu_substep(cuda_dev, spec, p, W, alpha)
    ref = _oracle_substep(spec, p, W, alpha)
    errs = {k: rel_err(got[k], ref[k]) for k in ("stress", "x", "v", "C", "F")}
    record("substep_fwd[block-1m-256]", **errs)
    for k, e in errs.items():
        assert e < 1e-4, (k, errs)


def test_raster_800_50k_fwd_bwd(cuda_dev):
    """config 3's render: 50 k Gaussians bound to the 100 k cloud, 800 x 800, sh 3, whole image."""
    from neuma_b200 import ops, synthetic as S
    spec = S.WORKLOADS["jellyduck-100k-128"]
    p = S.make_particles(spec, 42)
    gs = S.make_gaussians(p["x"], spec["K"], spec["G"], 43)
    cam = S.make_camera(800, 800)
    tan = math.tan(cam.FoVx * 0.5)
    cs = ops.raster_settings(800, 800, tan, tan, [1.0, 1.0, 1.0], 1.0, cam.world_view_transform,
                             cam.full_proj_transform, 3, cam.camera_center)
    s = RO.Settings(800, 800, tan, tan, np.ones(3, np.float32), cam.world_view_transform.numpy(),
                    cam.full_proj_transform.numpy(), 3, cam.camera_center.numpy())
    dev = cuda_dev
    means, shs, opac, cov = [gs[k].to(dev) for k in ("xyz", "shs", "opac", "cov6")]
    img, radii = ops.raster_fwd(cs, means, shs, None, opac, cov)
    G = torch.randn(3, 800, 800, generator=torch.Generator().manual_seed(1))
    g3 = ops.raster_bwd(cs, means, shs, None, opac, cov, radii, G.to(dev))[0]
    # ---- oracle, fp64
    b = RO.bin_np(s, gs["xyz"].numpy(), gs["cov6"].numpy())
    assert np.array_equal(radii.cpu().numpy(), b["radii"])
    keys, ids, ranges = ops.raster_binning(img, b["tiles_x"] * b["tiles_y"])
    assert np.array_equal(keys.cpu().numpy().view(np.uint64), b["keys"])              # tile / sort indices: bit-exact
    assert np.array_equal(ids.cpu().numpy(), b["ids"]) and np.array_equal(ranges.cpu().numpy(), b["ranges"])
    m = gs["xyz"].double().requires_grad_(True)
    ref, aux = RO.render_torch(s, b, m, gs["shs"].double(), None, gs["opac"].double(), gs["cov6"].double(),
                               return_aux=True)
    ok = ~aux["borderline"]
    diff = (img.cpu().double() - ref.detach()).abs().max(0).values
    og, = torch.autograd.grad(ref, m, G.double())
    e_l2 = float((g3.cpu().double() - og).norm() / og.norm())
    record("raster_800_50k", img_max_ok=float(diff[ok].max()), img_max_all=float(diff.max()),
           ok_frac=float(ok.float().mean()), g3_rel_l2=e_l2, g3_per_element=rel_err(g3, og),
           instances=float(keys.shape[0]))
    assert ok.float().mean() > 0.97
    assert float(diff[ok].max()) < 1e-4 * max(1.0, float(ref.abs().max()))
    assert float(diff.max()) < 2e-2                                  # a pixel may flip one 1/255-alpha contribution
    assert e_l2 < 5e-4
    # per element: a Gaussian whose alpha sits on the 1/255 / T < 1e-4 thresholds at some pixel takes
    # or loses that pixel's whole contribution (fp32 vs fp64 exp); 5e-3 measured on 50 k Gaussians
    assert rel_err(g3, og) < 2e-2
