"""GPU, end to end through the reference-shaped Python API (the drop-in classes):
finetune.py's hot loop -- elasticity -> MPMCacheDiffSim -> plasticity -> bindings ->
diff_rasterization -> l2 loss -> backward -- against the CPU oracle rollout."""
import types

import numpy as np
import pytest
import torch

from oracle import constitutive_oracle as CO
from oracle import mpm_oracle as MO
from oracle import pipeline_oracle as PO
from oracle import raster_oracle as RO
from scene import cov6_from_scale_rot
from test_bind_gpu import make_bindings
from test_constitutive_oracle import load_gold, weights_and_lora
from util import rel_err

pytestmark = pytest.mark.gpu


class FakeGaussians:
    """The four accessors diff_rasterization uses (modules/d3gs/scene/gaussian_model.py:95-118)."""

    def __init__(self, xyz, cov6, opac, shs, deg):
        self._xyz, self._cov, self.get_opacity, self.get_features = xyz, cov6, opac, shs
        self.active_sh_degree = deg

    @property
    def get_xyz(self):
        return self._xyz

    def get_covariance(self, scaling_modifier=1):
        return self._cov


def build_case(n=3000, K=1200, G=32, W=96, H=96, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = 0.35 + 0.3 * torch.rand(n, 3, generator=g)
    v = torch.tensor([0.0, -0.5, 0.0]).expand(n, 3).contiguous() + 0.05 * torch.randn(n, 3, generator=g)
    C = torch.zeros(n, 3, 3)
    F = torch.eye(3).expand(n, 3, 3).contiguous() + 0.03 * torch.randn(n, 3, 3, generator=g)
    vol = torch.full((n,), 0.3 ** 3 / n)
    rho = torch.full((n,), 1000.0)
    clip = torch.full((n,), 0.1)
    enabled = torch.ones(n, dtype=torch.int32)
    rowptr, col, w, B = make_bindings(K, n, seed=seed + 1)
    counts = (rowptr[1:] - rowptr[:-1])
    # every kernel bound to >= 1 particle (upstream prunes unbound kernels, tune/utils.py:157-320)
    assert int(counts.min()) >= 0
    k0 = 0.35 + 0.3 * torch.rand(K, 3, generator=g)
    scale = (1 / 48) * torch.exp(torch.empty(K, 3).uniform_(np.log(0.5), np.log(2.0), generator=g))
    cov6 = cov6_from_scale_rot(scale, torch.randn(K, 4, generator=g))
    opac = torch.empty(K, 1).uniform_(0.3, 0.95, generator=g)
    shs = 0.3 * torch.randn(K, 16, 3, generator=g)
    s = RO.look_at_camera(W=W, H=H)
    s.bg = np.ones(3, np.float32)
    gt = torch.rand(3, H, W, generator=g)
    return dict(x=x, v=v, C=C, F=F, vol=vol, rho=rho, clip=clip, enabled=enabled, rowptr=rowptr, col=col, w=w,
                B=B, k0=k0, cov6=cov6, opac=opac, shs=shs, s=s, gt=gt, n=n, K=K, G=G)


def camera_ns(s, dev):
    import math
    return types.SimpleNamespace(FoVx=2 * math.atan(s.tanfovx), FoVy=2 * math.atan(s.tanfovy),
                                 image_height=s.image_height, image_width=s.image_width,
                                 world_view_transform=torch.from_numpy(s.viewmatrix).to(dev),
                                 full_proj_transform=torch.from_numpy(s.projmatrix).to(dev),
                                 camera_center=torch.from_numpy(s.campos).to(dev))


@pytest.mark.parametrize("substeps,frames", [(1, 2), (3, 1)])
def test_finetune_loop_gradients(cuda_dev, substeps, frames):
    from neuma_b200.sim import MPMModelBuilder, MPMCacheDiffSim, MPMStaticsInitializer, MPMInitData
    from neuma_b200.material import InvariantFullMetaElasticity, InvariantFullMetaPlasticity
    from neuma_b200.glue import compute_bindings_xyz, compute_bindings_F, diff_rasterization, l2_loss, Bindings
    dev = cuda_dev
    cse = build_case()
    gold = load_gold("jelly")
    We, lora_e = weights_and_lora(gold, "elastic")
    Wp, lora_p = weights_and_lora(gold, "plastic")
    alpha = float(gold["alpha"])
    # ---- reference-shaped setup (finetune.py:266-313)
    sim_cfg = dict(num_grids=cse["G"], dt=1e-3, bound=1, gravity=[0.0, -9.8, 0.0], bc="noslip", eps=6e-7)
    model = MPMModelBuilder().parse_cfg(sim_cfg).finalize(dev, requires_grad=True)
    nsteps = substeps * frames
    sim = MPMCacheDiffSim(model, nsteps)
    init = MPMInitData(rho=1000.0, clip_bound=0.1, span=(0, nsteps), num_particles=cse["n"],
                       vol=float(cse["vol"][0]), pos=cse["x"].numpy())
    si = MPMStaticsInitializer(model)
    si.add_group(init)
    statics = si.finalize()
    mcfg = dict(layer_widths=[64, 64], norm=None, nonlinearity="gelu", no_bias=True, normalize_input=True, alpha=alpha)
    elasticity = InvariantFullMetaElasticity(mcfg).to(dev)
    plasticity = InvariantFullMetaPlasticity(mcfg).to(dev)
    for mod, tag in ((elasticity, "elastic"), (plasticity, "plastic")):
        mod.init_lora_layers(r=16, lora_alpha=16)
        sd = {k[len(f"param_{tag}."):]: v.float() for k, v in gold.items() if k.startswith(f"param_{tag}.")}
        mod.load_state_dict(sd)          # same parameter names as the reference checkpoints
        mod.to(dev)
        mod.freeze_all_except_lora()
    gaussians = FakeGaussians(cse["k0"].to(dev), cse["cov6"].to(dev), cse["opac"].to(dev), cse["shs"].to(dev), 3)
    cam = camera_ns(cse["s"], dev)
    bg = torch.ones(3, device=dev)
    bindings = torch.sparse_coo_tensor(cse["B"].indices(), cse["B"].values().float(), cse["B"].shape).to(dev)
    center, size = torch.tensor([0.1, 0.0, 0.1], device=dev), torch.tensor([0.8, 0.8, 0.8], device=dev)
    gts = [cse["gt"], cse["gt"].flip(2)][:frames]
    # ---- the hot loop, verbatim structure of finetune.py:334-414
    x, v, C, F = [cse[k].to(dev) for k in ("x", "v", "C", "F")]
    de_x_prev = ((x - center) / size).clone().detach()
    g_prev = gaussians.get_xyz.clone().detach()
    loss_rgb = 0.0
    images = []
    for it in range(nsteps):
        stress = elasticity(F)
        x, v, C, F = sim(statics, it, x, v, C, F, stress)
        F = plasticity(F)
        if (it + 1) % substeps == 0:
            de_x = (x.clone() - center) / size
            means3D = compute_bindings_xyz(de_x, de_x_prev, g_prev, bindings)
            deform_grad = compute_bindings_F(F, bindings)
            render = diff_rasterization(means3D, deform_grad, gaussians, cam, bg)
            gt = gts[(it + 1) // substeps - 1].to(dev)
            loss_rgb = loss_rgb + l2_loss(render, gt)
            images.append(render.detach())
            de_x_prev = de_x.clone().detach()
            g_prev = means3D.clone().detach()
    loss_rgb.backward()
    # ---- oracle
    mk = lambda lo: [(A.clone().requires_grad_(True), B_.clone().requires_grad_(True), s_) for A, B_, s_ in lo]
    lo_e, lo_p = mk(lora_e), mk(lora_p)
    co = MO.MPMConst(num_grids=cse["G"], dt=1e-3, bound=1, eps=6e-7, bc="noslip")
    oloss, oimgs, _ = PO.rollout_loss(co, cse["x"], cse["v"], cse["C"], cse["F"], cse["vol"], cse["rho"], cse["clip"],
                                      cse["enabled"], We, lo_e, Wp, lo_p, alpha, cse["B"], cse["k0"], cse["cov6"],
                                      cse["opac"], cse["shs"], cse["s"], gts, substeps,
                                      center.cpu(), size.cpu(), return_images=True)
    leaves = [t for lo in (lo_e, lo_p) for A, B_, _ in lo for t in (A, B_)]
    og = torch.autograd.grad(oloss, leaves)
    assert abs(float(loss_rgb) - float(oloss)) / float(oloss) < 1e-4
    for a, b in zip(images, oimgs):
        assert float((a.cpu().double() - b).abs().mean()) < 1e-4
    names = ["layers.0", "layers.1", "final_layer"]
    i = 0
    errs = {}
    for mod in (elasticity, plasticity):
        sd = dict(mod.named_parameters())
        for l in names:
            for ab in ("lora_A", "lora_B"):
                g = sd[f"{l}.fc.{ab}"].grad
                ref = og[i]
                i += 1
                denom = float(ref.norm()) + 1e-30
                errs[f"{type(mod).__name__[17:]}.{l}.{ab}"] = float((g.cpu().double() - ref).norm()) / denom
    print("relative L2 error of LoRA gradients:", {k: f"{v:.2e}" for k, v in errs.items()})
    # fp32 device rollout (atomics, rasterizer threshold flips) against an fp64 oracle
    assert max(errs.values()) < 2e-2, errs


def test_forward_sim_render_loop(cuda_dev):
    """render.py:304-332: MPMForwardSim in place + update_enabled + render, no grad."""
    from neuma_b200.sim import MPMModelBuilder, MPMForwardSim, MPMStateInitializer, MPMStaticsInitializer, MPMInitData
    from neuma_b200.material import InvariantFullMetaElasticity, InvariantFullMetaPlasticity
    from neuma_b200.glue import compute_bindings_xyz, compute_bindings_F, diff_rasterization
    dev = cuda_dev
    cse = build_case(n=2000, K=800, seed=3)
    gold = load_gold("plasticine")
    We, _ = weights_and_lora(gold, "elastic")
    Wp, _ = weights_and_lora(gold, "plastic")
    alpha = float(gold["alpha"])
    sim_cfg = types.SimpleNamespace(num_grids=cse["G"], dt=5e-4, bound=1, gravity=[0.0, -9.8, 0.0], bc="noslip", eps=0.0)
    model = MPMModelBuilder().parse_cfg(sim_cfg).finalize(dev, False)
    sim = MPMForwardSim(model)
    init = MPMInitData(rho=1000.0, clip_bound=0.1, span=(0, 4), num_particles=cse["n"], vol=float(cse["vol"][0]),
                       pos=cse["x"].numpy(), lin_vel=np.array([0.0, -0.5, 0.0]))
    sti, si = MPMStateInitializer(model), MPMStaticsInitializer(model)
    sti.add_group(init)
    si.add_group(init)
    state, _ = sti.finalize()
    statics = si.finalize()
    mcfg = dict(layer_widths=[64, 64], norm=None, nonlinearity="gelu", no_bias=True, normalize_input=True, alpha=alpha)
    elasticity = InvariantFullMetaElasticity(mcfg)
    plasticity = InvariantFullMetaPlasticity(mcfg)
    elasticity.load_state_dict({"layers.0.fc.weight": We[0].float(), "layers.1.fc.weight": We[1].float(),
                                "final_layer.fc.weight": We[2].float()})
    plasticity.load_state_dict({"layers.0.fc.weight": Wp[0].float(), "layers.1.fc.weight": Wp[1].float(),
                                "final_layer.fc.weight": Wp[2].float()})
    elasticity.to(dev).eval(); plasticity.to(dev).eval()
    gaussians = FakeGaussians(cse["k0"].to(dev), cse["cov6"].to(dev), cse["opac"].to(dev), cse["shs"].to(dev), 3)
    cam, bg = camera_ns(cse["s"], dev), torch.ones(3, device=dev)
    bindings = torch.sparse_coo_tensor(cse["B"].indices(), cse["B"].values().float(), cse["B"].shape).to(dev)
    center, size = torch.tensor([0.1, 0.0, 0.1], device=dev), torch.tensor([0.8, 0.8, 0.8], device=dev)
    with torch.no_grad():
        x, v, C, F, _ = state.to_torch()
        F = F.clone()
        state.from_torch(x=x, v=v, C=C, F=F)
        de_x_prev = ((x - center) / size).clone()
        g_prev = gaussians.get_xyz.clone()
        for step in range(1, 4):
            stress = elasticity(F)
            state.from_torch(stress=stress)
            x, v, C, F = sim(statics, state)
            F = plasticity(F)
            state.from_torch(F=F)
            si.update(statics, step)
            de_x = (x.clone() - center) / size
            means3D = compute_bindings_xyz(de_x, de_x_prev, g_prev, bindings)
            deform_grad = compute_bindings_F(F, bindings)
            render = diff_rasterization(means3D, deform_grad, gaussians, cam, bg)
            de_x_prev, g_prev = de_x.clone(), means3D.clone()
    co = MO.MPMConst(num_grids=cse["G"], dt=5e-4, bound=1, eps=0.0, bc="noslip")
    v0 = torch.tensor([0.0, -0.5, 0.0]).expand(cse["n"], 3)
    F0 = torch.eye(3).expand(cse["n"], 3, 3)
    _, oimgs, (ox, ov, oC, oF) = PO.rollout_loss(
        co, cse["x"], v0, cse["C"], F0, cse["vol"], cse["rho"], cse["clip"], cse["enabled"], We, None, Wp, None,
        alpha, cse["B"], cse["k0"], cse["cov6"], cse["opac"], cse["shs"], cse["s"], [cse["gt"]] * 3, 1,
        center.cpu(), size.cpu(), return_images=True)
    assert rel_err(x, ox) < 1e-4 and rel_err(v, ov) < 1e-4 and rel_err(F, oF) < 1e-4
    assert float((render.cpu().double() - oimgs[-1]).abs().mean()) < 1e-4
