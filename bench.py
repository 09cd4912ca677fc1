#!/usr/bin/env python
"""bench.py -- sim+render fwd+bwd frames/s at 100k particles / 128^3 grid (BASELINE.json).

One "step" = one rendered frame of finetune.py's loop (experiments/finetune.py:360-414):
S MPM substeps (elasticity -> P2G/grid/G2P -> plasticity), particle->Gaussian binding,
covariance deformation, one 800x800 rasterised view, l2 loss against a fixed GT image,
and the complete reverse pass down to dL/dW (LoRA-merged weights) and dL/d(state at
frame start).  Workload: SURVEY 8d config 3 ("jellyduck-100k-128"), synthetic.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--substeps S] [--impl reference]

Prints ONE JSON line (rank 0).  N > 1: one process per GPU (torchrun), every rank runs
its own frame stream (the render path replicates; see DESIGN.md "multi-GPU"), weak scaling.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np
import torch

WORKLOAD = "jellyduck-100k-128"
IMAGE = 800
GOLD = os.path.join(ROOT, "tests", "golden", "material_jelly.npz")


def load_weights():
    """Shipped jelly_0300 base weights + a rank-16 LoRA adaptor (fixture made from the
    reference's own classes, tests/golden/make_golden_material.py), merged: W + s B A."""
    z = np.load(GOLD)
    out = {}
    for tag in ("elastic", "plastic"):
        Ws = []
        for l in ("layers.0", "layers.1", "final_layer"):
            W = torch.from_numpy(z[f"param_{tag}.{l}.fc.weight"]).float()
            A = torch.from_numpy(z[f"param_{tag}.{l}.fc.lora_A"]).float()
            B = torch.from_numpy(z[f"param_{tag}.{l}.fc.lora_B"]).float()
            Ws.append(W + float(z["lora_scaling"]) * (B @ A))
        out[tag] = Ws
    out["alpha"] = float(z["alpha"])
    return out


def build_workload(substeps: int, seed: int = 42):
    from neuma_b200 import synthetic as S
    torch.manual_seed(seed)                      # finetune.py:496-499
    spec = S.WORKLOADS[WORKLOAD]
    p = S.make_particles(spec, seed)
    g = S.make_gaussians(p["x"], spec["K"], spec["G"], seed + 1)
    cam = S.make_camera(IMAGE, IMAGE)
    gt = torch.rand(3, IMAGE, IMAGE, generator=torch.Generator().manual_seed(seed + 2))
    return dict(spec=spec, p=p, g=g, cam=cam, gt=gt, w=load_weights(), S=substeps)


def workload_config(wl, world: int) -> dict:
    """The `config` object of the JSON line -- ONE function for both arms, so that the driver's
    same_config check compares like with like."""
    spec, g = wl["spec"], wl["g"]
    return {"workload": WORKLOAD, "particles": spec["n"], "grid": spec["G"], "gaussians": g["K"],
            "image": [IMAGE, IMAGE], "sh_degree": 3, "substeps_per_frame": wl["S"], "views": 1, "dt": spec["dt"],
            "mode": "fwd+bwd (l2 loss -> dL/dW_eff, dL/dstate0)",
            "weights": os.path.basename(GOLD)[9:-4] + "_0300 + LoRA r16 merged",
            "l2": "flushed between timed steps (256 MiB write)", "parallelism": f"replicas x{world}"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 6 for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def event_time_us(fn, iters, warm, flush=None):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return float(np.mean(ts))


def run_gpu(args):
    import torch.distributed as dist
    from neuma_b200 import ops, build
    from neuma_b200.pipeline import FusedFrame
    from neuma_b200.glue import Bindings
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    build.build()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    wl = build_workload(args.substeps)
    spec, p, g, S = wl["spec"], wl["p"], wl["g"], wl["S"]
    n, G, K = spec["n"], spec["G"], g["K"]
    cc = ops.mpm_const(G, spec["dt"], 1, (0.0, -9.8, 0.0), 6e-7, "noslip")
    d = {k: v.to(dev) for k, v in p.items()}
    statics = (d["vol"], d["rho"], d["clip"], d["enabled"])
    We = [w.to(dev) for w in wl["w"]["elastic"]]
    Wp = [w.to(dev) for w in wl["w"]["plastic"]]
    bind = Bindings(g["rowptr"].to(dev), g["col"].to(dev), g["w"].to(dev), n)
    gs = {k: g[k].to(dev) for k in ("cov6", "opac", "shs")}
    cam = wl["cam"]
    import math
    settings = ops.raster_settings(IMAGE, IMAGE, math.tan(cam.FoVx * 0.5), math.tan(cam.FoVy * 0.5),
                                   [1.0, 1.0, 1.0], 1.0, cam.world_view_transform, cam.full_proj_transform, 3,
                                   cam.camera_center)
    # warm-up rollout: 10 substeps so that F is off identity
    pre = FusedFrame(n, cc, statics, We, Wp, wl["w"]["alpha"], 10, dev, use_graphs=False)
    pre.load_state(d["x"], d["v"], d["C"], d["F"])
    x0, v0, C0, F0 = [t.clone() for t in pre.sim_forward()]
    del pre
    eng = FusedFrame(n, cc, statics, We, Wp, wl["w"]["alpha"], S, dev, bind=bind, gaussians=gs, settings=settings,
                     fuse=args.fuse, save_svd=args.save_svd)
    eng.load_state(x0, v0, C0, F0)
    k_prev = g["xyz"].to(dev)
    de_x_prev = d["x"].clone()
    gt = wl["gt"].to(dev)
    ctx = ops._ctx(x0)

    # launches per frame: count one eager (un-graphed) frame with the library's own counter
    eager = FusedFrame(n, cc, statics, We, Wp, wl["w"]["alpha"], S, dev, bind=bind, gaussians=gs,
                       settings=settings, use_graphs=False, fuse=args.fuse, save_svd=args.save_svd)
    eager.load_state(x0, v0, C0, F0)
    l0 = ctx.launches
    eager.frame_fwd_bwd(gt, de_x_prev, k_prev)
    torch.cuda.synchronize()
    launches_per_frame = ctx.launches - l0
    del eager

    def step():
        eng.zero_weight_grads()
        return eng.frame_fwd_bwd(gt, de_x_prev, k_prev)

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    total_ms = 0.0
    with ClockSampler(local) as clk:
        for _ in range(args.steps):
            flush.zero_()                                    # L2 flush between timed iterations
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            loss, g0 = step()
            b.record()
            b.synchronize()
            total_ms += a.elapsed_time(b)
    torch.cuda.synchronize()
    t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.barrier()
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / args.steps
    value = world * 1000.0 / ms_per_step

    # ---- e2e: same frame through the engine with HOST buffers (pinned), copies inside the timed region
    hx, hv, hC, hF = [t_.cpu().pin_memory() for t_ in (x0, v0, C0, F0)]
    hgt = wl["gt"].pin_memory()
    hloss = torch.empty((), dtype=torch.float32).pin_memory()
    hgw = [torch.empty_like(w, device="cpu").pin_memory() for w in eng.gWe + eng.gWp]
    dgt = torch.empty_like(gt)

    copy_stream = torch.cuda.Stream(dev)
    gt_ready = torch.cuda.Event()

    def step_e2e():
        # the GT image is not needed before the loss: its H2D copy runs on a copy stream under the
        # simulation (finetune.py:386 does it synchronously per frame); the state copy cannot overlap
        eng.load_state(hx, hv, hC, hF)
        copy_stream.wait_stream(torch.cuda.current_stream(dev))   # after the state copy: it owns the link first
        with torch.cuda.stream(copy_stream):
            dgt.copy_(hgt, non_blocking=True)
            gt_ready.record(copy_stream)
        eng.zero_weight_grads()
        loss_, _ = eng.frame_fwd_bwd(dgt, de_x_prev, k_prev, gt_ready=gt_ready)
        hloss.copy_(loss_, non_blocking=True)
        for h_, w_ in zip(hgw, eng.gWe + eng.gWp):
            h_.copy_(w_, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    for _ in range(3):
        step_e2e()
    torch.cuda.synchronize()
    e2e_ms = 0.0
    for _ in range(args.steps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        step_e2e()
        b.record()
        b.synchronize()
        e2e_ms += a.elapsed_time(b)
    t = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms_per_step = float(t.item()) / args.steps
    h2d = 4 * (hx.numel() + hv.numel() + hC.numel() + hF.numel() + hgt.numel())
    d2h = 4 * (1 + sum(h_.numel() for h_ in hgw))

    # ---- BASELINE config 5 in the same invocation: the 1M / 256^3 rollout sharded by Z-slab over the
    # ranks (strong scaling, one halo exchange per substep each way) -- the multi-GPU path of the
    # hot path; the headline metric above stays the 100k frame (replicas: the render does not shard)
    slab = None
    if not args.no_slab:
        del flush
        torch.cuda.empty_cache()
        slab = slab_measure(dev, world, rank, S, args.slab_steps, 3)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- per-kernel durations at the workload's shapes (CUDA events, L2 flushed), for the roofline
    Ng = G ** 3
    s0 = eng.S - 1
    xs, vs, Cs, Fs, st = eng.x[s0], eng.v[s0], eng.C[s0], eng.F[s0], eng.stress[s0]
    outb = tuple(torch.empty_like(t_) for t_ in (xs, vs, Cs, Fs))
    gout = [torch.randn_like(t_) for t_ in outb]
    gF_ = torch.empty_like(Fs)
    part_tmp = ops.constitutive_partials(Fs)
    gouts5 = [torch.empty_like(t_) for t_ in (xs, vs, Cs, Fs, st)]
    it, wm = 20, 3
    kt = {}
    sx = (eng.vol, eng.rho, eng.clip, eng.enabled)                 # statics in the engine's cell order
    ce = eng.c                                                     # the engine's const: compact grid slots
    grid_s, adj_s, box_s = eng.grid[s0], eng.adj, eng.box[s0]
    box_tmp = torch.zeros(8, dtype=torch.int32, device=dev)
    kt["grid_clear"] = event_time_us(lambda: ops.mpm_box_clear(ce, box_s, grid_s, None, box_tmp), it, wm, flush)
    kt["p2g"] = event_time_us(lambda: ops.mpm_p2g_to(ce, xs, vs, Cs, st, sx[0], sx[1], sx[3], grid_s, box=box_s),
                              it, wm, flush)
    def _clear_p2g():
        ops.mpm_box_clear(ce, box_s, grid_s, None, box_tmp)
        ops.mpm_p2g_to(ce, xs, vs, Cs, st, sx[0], sx[1], sx[3], grid_s, box=box_s)
    kt["clear+p2g"] = event_time_us(_clear_p2g, it, wm, flush)          # one timed region: the metric's t(clear + P2G)

    def _mpm_fwd():
        ops.mpm_rollout_fwd(ce, xs, vs, Cs, Fs, st, *sx, outb, grid_s, None, box_s, box_tmp)
    kt["mpm_substep_fwd"] = event_time_us(_mpm_fwd, it, wm, flush)
    kt["grid_op"] = event_time_us(lambda: ops.mpm_grid_op_box(ce, box_s, grid_s, None), it, wm, flush)
    kt["g2p"] = event_time_us(lambda: ops.mpm_g2p_from(ce, xs, Fs, sx[2], sx[3], grid_s, outb, box_tmp, box=box_s),
                              it, wm, flush)
    # leave a valid (v, m) grid in the slot for the backward timing
    ops.mpm_rollout_fwd(ce, xs, vs, Cs, Fs, st, *sx, outb, grid_s, None, box_s, box_tmp)

    def _bwd():                       # as the engine runs it: clear the adjoint target over the box, then the 3 passes
        ops.mpm_box_clear(ce, box_s, adj_s, None, None)
        ops.mpm_rollout_bwd(ce, xs, vs, Cs, Fs, st, *sx, outb[1], outb[2], *gout, grid_s, adj_s, box_s, *gouts5)
    kt["mpm_substep_bwd"] = event_time_us(_bwd, it, wm, flush)
    kt["constitutive_fwd"] = event_time_us(lambda: ops.constitutive_fwd_staged(0, Fs, eng.We_st, 0.0, gF_), it, wm, flush)
    kt["constitutive_bwd"] = event_time_us(
        lambda: ops.constitutive_bwd_partials(0, Fs, eng.We_st, 0.0, gout[3], gF_, part_tmp), it, wm, flush)
    rs = eng._render_saved
    kt["raster_fwd"] = event_time_us(lambda: ops.raster_fwd(settings, rs[0], gs["shs"], None, gs["opac"], rs[1]), 10, 2, flush)
    dimg = torch.randn(3, IMAGE, IMAGE, device=dev)
    kt["raster_bwd"] = event_time_us(
        lambda: ops.raster_bwd(settings, rs[0], gs["shs"], None, gs["opac"], rs[1], rs[2], dimg), 10, 2, flush)
    # in-situ breakdown of the frame (CUDA-graph replays / render calls, events, no flush)
    def _ms(fn, reps=5):
        fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        b.synchronize()
        return a.elapsed_time(b) / reps
    gz = [torch.zeros_like(t_) for t_ in (x0, v0, C0, F0)]
    frame_ms = {"sim_fwd_graph": _ms(lambda: eng.sim_forward()),
                "sim_bwd_graph": _ms(lambda: eng.sim_backward(*gz)),
                "render_fwd": _ms(lambda: eng.render_forward(de_x_prev, k_prev)),
                "render_bwd": _ms(lambda: eng.render_backward(dimg, accumulate_into_state_grad=True))}
    nr = ops.raster_binning(x0, 50 * 50)[0].shape[0]
    per_frame_us = {"constitutive_bwd": 2 * S * kt["constitutive_bwd"], "constitutive_fwd": 2 * S * kt["constitutive_fwd"],
                    "mpm_fwd(clear+p2g+grid_op+g2p)": S * kt["mpm_substep_fwd"],
                    "mpm_substep_bwd": S * kt["mpm_substep_bwd"], "raster_fwd": kt["raster_fwd"],
                    "raster_bwd": kt["raster_bwd"]}
    peaks = {}
    pk_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk_path):
        peaks = json.load(open(pk_path))
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    tf_peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)"
    p2g_bytes = 108 * n + 44 * Ng                              # clear 28 Ng + P2G 108 Np + 16 Ng
    p2g_gbs = p2g_bytes / kt["clear+p2g"] / 1e3
    fwd_bytes = 260 * n + 84 * Ng
    mpm_fwd_gbs = fwd_bytes / kt["mpm_substep_fwd"] / 1e3
    dominant = max(per_frame_us, key=per_frame_us.get)
    traffic = {}
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath))
    if dominant.startswith("constitutive"):
        # fused SVD + 13-64-64-9 MLP (+ LoRA merged); bwd = recompute + dgrad + wgrad = 3 x 5504 MAC / particle
        macs = 5504 * (3 if dominant.endswith("bwd") else 1)
        achieved = 2.0 * macs * n / (kt[dominant] * 1e-6) / 1e12
        roof = {"kernel": dominant + "_kernel", "bound": "tensor", "achieved": achieved, "peak": tf_peak,
                "unit": "TFLOP/s", "frac": achieved / tf_peak,
                "traffic": traffic.get(dominant + "_kernel", {}).get("dram_bytes_per_launch"),
                "note": "tcgen05 kind::f16, fp32 split into bf16 terms (forward x3 = 6 products, backward x2 = 3 products per "
                        "fp32 MAC, counted as ONE MAC here); "
                        "SVD/GELU/split CUDA-core work bounds the kernel, not the tensor pipe; peak = sustained bf16 "
                        + peak_src}
    else:
        roof = {"kernel": "grid_clear+p2g", "bound": "hbm", "achieved": p2g_gbs, "peak": hbm_peak, "unit": "GB/s",
                "frac": p2g_gbs / hbm_peak, "traffic": traffic.get("box_clear+p2g_to", {}).get("dram_bytes_per_launch"), "note": "dense-grid algorithmic bytes (SURVEY 8d) over a box-sparse clear + P2G: the real traffic is "
                        "smaller, so the fraction can exceed 1; peak " + peak_src}
    line = {
        "metric": "sim+render fwd+bwd frames/sec at 100k particles/128^3 grid; P2G HBM GB/s",
        "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(wl, world),
        "instances_rendered": int(nr),
        "rollout_memory": {**eng.memory_bytes_per_substep(), "reference_per_particle": 528,
                           "note": "bytes one substep keeps resident for the backward: compact particle state "
                                   "(x,v,C,F,stress,F_trial) + one box-sized saved grid; the reference's "
                                   "MPMCacheDiffSim pins 528 B/particle/substep (interface.py:108-123)"},
        "p2g_hbm_gbs": p2g_gbs, "p2g_hbm_frac_of_measured_peak": p2g_gbs / hbm_peak,
        "p2g_hbm_frac_of_nominal_8TBs": p2g_gbs / 8000.0,
        "mpm_fwd_substep_gbs": mpm_fwd_gbs,
        "kernel_us": {k: round(v, 2) for k, v in kt.items()},
        "frame_share_us": {k: round(v, 1) for k, v in per_frame_us.items()},
        "frame_ms_in_situ": {k: round(v, 3) for k, v in frame_ms.items()},
        "roofline": roof,
        "e2e": {"value": world * 1000.0 / e2e_ms_per_step, "unit": "frames/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms_per_step},
        "gpu_launches": int(launches_per_frame * args.steps),
        "gpu_launches_per_frame": int(launches_per_frame),
        "clocks": clk.summary(),
        "loss": float(loss),
    }
    if slab is not None:
        line["slab"] = slab
    if not args.no_dropin:
        line["dropin"] = dropin_measure(wl, dev, (x0, v0, C0, F0))
        line["dropin_frames_per_s"] = line["dropin"]["frames_per_s"]
        # + ONE extra line in the caller: `elasticity.cache_merged = plasticity.cache_merged = True`
        line["dropin_cached_merge_frames_per_s"] = dropin_measure(wl, dev, (x0, v0, C0, F0),
                                                                   cache_merged=True)["frames_per_s"]
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_sample(wl, x0.cpu(), v0.cpu(), C0.cpu(), F0.cpu())
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# --------------------------------------------------------------------------- drop-in (Level-1) path
def dropin_measure(wl, dev, state0, steps: int = 5, warmup: int = 2, cache_merged: bool = False) -> dict:
    """The SAME frame through the reference-shaped classes with autograd, exactly as finetune.py's hot
    loop calls them after the import swap of INTEGRATION.md (elasticity -> MPMCacheDiffSim ->
    plasticity per substep, compute_bindings_*, diff_rasterization, l2_loss, loss.backward()): what a
    user gets WITHOUT adopting FusedFrame."""
    import types
    from neuma_b200.sim import MPMModelBuilder, MPMCacheDiffSim, MPMStaticsInitializer, MPMInitData
    from neuma_b200.material import InvariantFullMetaElasticity, InvariantFullMetaPlasticity
    from neuma_b200.glue import compute_bindings_xyz, compute_bindings_F, diff_rasterization, l2_loss
    spec, p, g, S = wl["spec"], wl["p"], wl["g"], wl["S"]
    z = np.load(GOLD)
    sim_cfg = dict(num_grids=spec["G"], dt=spec["dt"], bound=1, gravity=[0.0, -9.8, 0.0], bc="noslip", eps=6e-7)
    model = MPMModelBuilder().parse_cfg(sim_cfg).finalize(dev, requires_grad=True)
    sim = MPMCacheDiffSim(model, S)
    init = MPMInitData(rho=1000.0, clip_bound=0.1, span=(0, S), num_particles=spec["n"], vol=float(p["vol"][0]),
                       pos=p["x"].numpy())
    si = MPMStaticsInitializer(model)
    si.add_group(init)
    statics = si.finalize()
    mcfg = dict(layer_widths=[64, 64], norm=None, nonlinearity="gelu", no_bias=True, normalize_input=True,
                alpha=float(z["alpha"]))
    mods = []
    for cls, tag in ((InvariantFullMetaElasticity, "elastic"), (InvariantFullMetaPlasticity, "plastic")):
        m = cls(mcfg).to(dev)
        m.init_lora_layers(r=16, lora_alpha=16)
        m.load_state_dict({k[len(f"param_{tag}."):]: torch.from_numpy(z[k]).float() for k in z.files
                           if k.startswith(f"param_{tag}.")})
        m.to(dev)
        m.freeze_all_except_lora()
        m.cache_merged = cache_merged      # opt-in: one shared LoRA-merge node per rollout (material.py)
        mods.append(m)
    elasticity, plasticity = mods

    class G_:                                   # the accessors diff_rasterization uses (gaussian_model.py:95-118)
        active_sh_degree = 3
        get_xyz = g["xyz"].to(dev)
        get_opacity = g["opac"].to(dev)
        get_features = g["shs"].to(dev)
        _cov = g["cov6"].to(dev)

        def get_covariance(self, scaling_modifier=1):
            return self._cov
    gaussians = G_()
    cam_ = wl["cam"]
    cam = types.SimpleNamespace(FoVx=cam_.FoVx, FoVy=cam_.FoVy, image_height=IMAGE, image_width=IMAGE,
                                world_view_transform=cam_.world_view_transform.to(dev),
                                full_proj_transform=cam_.full_proj_transform.to(dev),
                                camera_center=cam_.camera_center.to(dev))
    bg = torch.ones(3, device=dev)
    B = torch.sparse_csr_tensor(g["rowptr"].long(), g["col"].long(), g["w"], (g["K"], spec["n"])).to_sparse_coo().to(dev)
    gt = wl["gt"].to(dev)
    x0, v0, C0, F0 = state0

    def frame():
        for m in mods:
            m.zero_grad(set_to_none=True)
        x, v, C, F = x0, v0, C0, F0
        de_x_prev = p["x"].to(dev)
        for it in range(S):
            stress = elasticity(F)
            x, v, C, F = sim(statics, it, x, v, C, F, stress)
            F = plasticity(F)
        means3D = compute_bindings_xyz(x, de_x_prev, gaussians.get_xyz, B)
        deform_grad = compute_bindings_F(F, B)
        render = diff_rasterization(means3D, deform_grad, gaussians, cam, bg)
        loss = l2_loss(render, gt)
        loss.backward()
        return loss
    for _ in range(warmup):
        frame()
    torch.cuda.synchronize()
    # this path is bound by the HOST (Python / autograd launch overhead, ~15 ms per frame): three timed blocks,
    # the best one is reported and all three are listed -- a busy host core shows up as a slow block
    blocks = []
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            loss = frame()
        b.record()
        b.synchronize()
        blocks.append(a.elapsed_time(b) / steps)
    ms = min(blocks)
    model.check(sim.curr_states)
    return {"frames_per_s": 1000.0 / ms, "ms_per_frame": ms, "ms_per_frame_blocks": blocks,
            "loss": float(loss.detach()), "steps": steps,
            "api": "neuma_b200.sim.MPMCacheDiffSim + material.InvariantFullMeta* (LoRA r16, autograd) + glue."
                   "diff_rasterization: the reference's finetune.py loop with the import swap only"}


# --------------------------------------------------------------------------- Z-slab mode
SLAB_WORKLOAD = "block-1m-256"
SLAB_FUSE_GATHER = False


def slab_measure(dev, world: int, rank: int, substeps: int, steps: int, warmup: int, halo: str = "fused",
                 variants: bool = True) -> dict:
    """BASELINE config 5: 1M particles / 256^3, one S-substep frame of MPM + both constitutive nets,
    forward + backward, particles and grid sharded by Z-slab over the ranks with ONE summing halo
    exchange per substep each way (neuma_b200/slab.py).  STRONG scaling: the cloud is fixed.
    Under torchrun rank 0 first times the same frame ALONE on its GPU (the N = 1 reference of the
    speed-up), then all ranks time the sharded frame; device-timed, max over ranks."""
    import torch.distributed as dist
    from neuma_b200 import synthetic as S
    from neuma_b200.slab import SlabRollout
    spec = S.WORKLOADS[SLAB_WORKLOAD]
    p = S.make_particles(spec, 42)
    w = load_weights()
    gen = torch.Generator(device="cpu").manual_seed(7)

    def make(standalone: bool, **kw):
        sim = SlabRollout(spec["G"], spec["dt"], w["elastic"], w["plastic"], w["alpha"], substeps, dev,
                          halo=halo, standalone=standalone, fuse_gather=SLAB_FUSE_GATHER, **kw)
        sim.distribute(p["x"], p["v"], p["C"], p["F"], p["vol"], p["rho"], p["clip"], p["enabled"])
        g = [torch.randn(sim.n, 3, generator=gen).to(dev), None, None,
             torch.randn(sim.n, 3, 3, generator=gen).to(dev)]
        return sim, g

    def time_ms(sim, g, sync_ranks: bool) -> float:
        def step():
            sim.forward()
            sim.backward(*g)
        for _ in range(max(warmup, 3)):
            step()
        torch.cuda.synchronize()
        if sync_ranks:
            dist.barrier()
            torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            step()
        b.record()
        b.synchronize()
        t = torch.tensor([a.elapsed_time(b)], device=dev, dtype=torch.float64)
        if sync_ranks:
            dist.barrier()
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / steps

    out = {"workload": SLAB_WORKLOAD, "particles": spec["n"], "grid": spec["G"], "substeps_per_frame": substeps,
           "dt": spec["dt"], "nranks": world, "scaling": "strong", "steps": steps,
           "mode": "MPM + constitutive fwd+bwd of one S-substep frame"}
    # ---- N = 1 reference, rank 0 alone
    t1 = torch.zeros(1, device=dev, dtype=torch.float64)
    if rank == 0:
        sim1, g1 = make(True)
        t1[0] = time_ms(sim1, g1, False)
        del sim1, g1
        torch.cuda.empty_cache()
    if world > 1:
        dist.barrier()
        dist.broadcast(t1, 0)
    out["ms_per_frame_n1"] = float(t1.item())
    if world == 1:
        out["ms_per_frame"] = out["ms_per_frame_n1"]
        out["speedup_vs_n1"] = 1.0
        return out
    # ---- sharded
    sim, g = make(False)
    ms = time_ms(sim, g, True)
    sim.check_drift()
    cnt = torch.tensor([sim.n], device=dev, dtype=torch.int64)
    allc = [torch.zeros_like(cnt) for _ in range(world)]
    dist.all_gather(allc, cnt)
    ov = [sim.plan.overlap(r, True) for r in range(world - 1)]
    out.update({"ms_per_frame": ms, "speedup_vs_n1": out["ms_per_frame_n1"] / ms, "frames_per_s": 1000.0 / ms,
                "halo": halo, "overlap": sim.overlap, "halo_k": sim.k, "cuda_graphs": sim.use_graphs,
                "slab_bounds": sim.plan.b, "particles_per_rank": [int(c_.item()) for c_ in allc],
                "exchanges_per_frame": 2 * substeps,
                "halo_bytes_dense_per_frame": sum(o[1] for o in ov) * spec["G"] ** 2 * 16 * 2 * 2 * substeps})
    if variants:
        # where the exchange time goes: the same frame with the other overlap modes, and without
        # any exchange (wrong results, valid timing)
        by_mode = {sim.overlap: ms}
        for mode in ("none", "split"):
            if mode not in by_mode and sim._fused is not None:
                sim.overlap, sim._graphs = mode, {}
                by_mode[mode] = time_ms(sim, g, True)
        out["ms_per_frame_by_overlap"] = by_mode
        if "split" in by_mode:   # what the split LAUNCHES cost by themselves (no exchange kernel at all)
            sim.overlap, sim.skip_exchange, sim._graphs = "split", True, {}
            out["ms_per_frame_split_no_exchange"] = time_ms(sim, g, True)
        sim.overlap, sim.skip_exchange, sim._graphs = "none", True, {}
        ms0 = time_ms(sim, g, True)
        out["ms_per_frame_no_exchange"] = ms0
        out["halo_us_per_exchange"] = (by_mode.get("none", ms) - ms0) * 1000.0 / (2 * substeps)
        out["halo_exposed_us_per_exchange"] = (ms - ms0) * 1000.0 / (2 * substeps)
    return out


def run_slab(args):
    """--slab: only the Z-slab measurement (its own JSON line; not the headline metric)."""
    import torch.distributed as dist
    from neuma_b200 import build
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    build.build()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    r = slab_measure(dev, world, rank, args.substeps, args.steps, args.warmup, args.halo)
    if rank == 0:
        print(json.dumps({
            "metric": "MPM+constitutive fwd+bwd frames/sec, Z-slab sharded", "value": 1000.0 / r["ms_per_frame"],
            "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": r["ms_per_frame"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": {"workload": SLAB_WORKLOAD,
                                                             "parallelism": f"z-slab x{world}"}, "slab": r}))
    if world > 1:
        dist.destroy_process_group()


# --------------------------------------------------------------------------- CPU oracle arm
def cpu_sample(wl, x0, v0, C0, F0, reps: int = 1):
    """Times the CPU oracle (a restatement of the reference's mpm.py / meta.py / the published
    rasterizer -- NOT Warp, which is not installable) on a bounded sample of the workload:
    ONE substep (elasticity -> sim -> plasticity) fwd+bwd and ONE render fwd+bwd at full
    size, then frames/s = 1 / (S * t_substep + t_render)."""
    from oracle import constitutive_oracle as CO, mpm_oracle as MO, raster_oracle as RO
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    spec, p, g, S = wl["spec"], wl["p"], wl["g"], wl["S"]
    co = MO.MPMConst(num_grids=spec["G"], dt=spec["dt"], bound=1, eps=6e-7, bc="noslip")
    We = [w.clone().requires_grad_(True) for w in wl["w"]["elastic"]]
    Wp = [w.clone().requires_grad_(True) for w in wl["w"]["plastic"]]
    t_sub = []
    for _ in range(reps):
        t0 = time.perf_counter()
        x = x0.clone().requires_grad_(True)
        F = F0.clone().requires_grad_(True)
        stress = CO.elasticity(F, We)
        x2, v2, C2, F2 = MO.substep_fwd(co, x, v0, C0, F, stress, p["vol"], p["rho"], p["clip"], p["enabled"])
        F3 = CO.plasticity(F2, Wp, wl["w"]["alpha"])
        (x2.sum() + v2.sum() + C2.sum() + F3.sum()).backward()
        t_sub.append(time.perf_counter() - t0)
    cam = wl["cam"]
    import math
    s = RO.Settings(IMAGE, IMAGE, math.tan(cam.FoVx * 0.5), math.tan(cam.FoVy * 0.5), np.ones(3, np.float32),
                    cam.world_view_transform.numpy(), cam.full_proj_transform.numpy(), 3, cam.camera_center.numpy())
    B = torch.sparse_csr_tensor(g["rowptr"].long(), g["col"].long(), g["w"], (g["K"], spec["n"])).to_sparse_coo()
    t0 = time.perf_counter()
    xx = x0.clone().requires_grad_(True)
    means3D = g["xyz"] + torch.sparse.mm(B, xx - p["x"])
    Fk = torch.sparse.mm(B, F0.reshape(-1, 9)).reshape(-1, 3, 3)
    cov = RO.deform_cov(g["cov6"], Fk).detach()
    binning = RO.bin_np(s, means3D.detach().numpy(), cov.numpy())
    img = RO.render_torch(s, binning, means3D, g["shs"], None, g["opac"], cov, dtype=torch.float32)
    ((img - wl["gt"]) ** 2).mean().backward()
    t_render = time.perf_counter() - t0
    ts = float(np.median(t_sub))
    try:                                   # the reference's own Warp-CPU path, if it were installed
        import warp  # noqa: F401
        warp_ok = True
    except Exception:
        warp_ok = False
    return {"value": 1.0 / (S * ts + t_render), "unit": "frames/s", "cores": cores, "kind": "port",
            "sampled": True, "warp_importable": warp_ok,
            "sample": f"1 of {S} substeps fwd+bwd ({ts:.2f} s) + 1 render 800x800 fwd+bwd ({t_render:.2f} s), "
                      f"torch-CPU fp32 oracle, extrapolated to S={S}",
            "t_substep_s": ts, "t_render_s": t_render}


def run_reference(args):
    """--impl reference: the reference's own CPU path.  warp-lang / diff_gaussian_rasterization are
    not installable (no network), so this times the oracle port on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    wl = build_workload(args.substeps)
    p = wl["p"]
    vals, secs = [], []
    base = None
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        base = cpu_sample(wl, p["x"], p["v"], p["C"], p["F"] + 0.01 * torch.randn_like(p["F"]))
        if i >= args.warmup:
            vals.append(base["value"])
            secs.append(time.perf_counter() - t0)
    v = float(np.mean(vals))
    line = {"impl": "reference", "metric": "sim+render fwd+bwd frames/sec at 100k particles/128^3 grid; P2G HBM GB/s",
            "value": v, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1000.0 / v, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": workload_config(wl, world),
            "cpu_baseline": {"value": v, "unit": "frames/s", "cores": base["cores"], "kind": "port",
                             "sample": base["sample"], "wall_s_per_sample": float(np.mean(secs))},
            "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    global WORKLOAD, GOLD, SLAB_WORKLOAD, SLAB_FUSE_GATHER
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--substeps", type=int, default=20, help="MPM substeps per rendered frame (S)")
    ap.add_argument("--impl", default="neuma_b200", choices=["neuma_b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default=WORKLOAD, choices=["bouncyball-50k-64", "jellyduck-100k-128", "claycat-200k-128"],
                    help="parity-test configs of BASELINE.json for the record; the headline is jellyduck-100k-128")
    ap.add_argument("--no-fuse", action="store_true",
                    help="A/B: every MPM stage its own kernel (same as --fuse none)")
    ap.add_argument("--fuse", default="epilogue", choices=["epilogue", "full", "backward-full", "none"],
                    help="A/B: epilogue (default) = P2G / the adjoint scatter ride in the constitutive kernels; full = also "
                         "G2P and the adjoint gather (3 kernels per substep each way); backward-full = epilogue forward, "
                         "full backward; none = every stage its own kernel")
    ap.add_argument("--save-svd", action="store_true",
                    help="A/B: the forward constitutive kernels save V, the backward ones skip the Jacobi iteration")
    ap.add_argument("--no-dropin", action="store_true", help="skip the drop-in (autograd classes) measurement")
    ap.add_argument("--no-slab", action="store_true", help="skip the Z-slab (config 5) section of the line")
    ap.add_argument("--slab-steps", type=int, default=5, help="timed frames of the Z-slab section")
    ap.add_argument("--slab-fuse-gather", action="store_true",
                    help="A/B: slab backward with the adjoint gather inside the elasticity backward's prologue")
    ap.add_argument("--slab-workload", default=SLAB_WORKLOAD, choices=["block-1m-256", "block-250k-256"],
                    help="block-250k-256: development proxy (2 GPUs carry an 8-GPU rank's load each)")
    ap.add_argument("--halo", default="fused", choices=["fused", "peer", "nccl"],
                    help="Z-slab halo exchange: one fused windowed kernel (default), round-1 push/add kernel pairs, "
                         "or NCCL send/recv (eager)")
    ap.add_argument("--slab", action="store_true",
                    help="BASELINE config 5: 1M particles / 256^3 MPM fwd+bwd, Z-slab sharded over the ranks")
    args = ap.parse_args()
    args.fuse = False if (args.no_fuse or args.fuse == "none") else args.fuse
    WORKLOAD = args.workload
    SLAB_WORKLOAD = args.slab_workload
    SLAB_FUSE_GATHER = args.slab_fuse_gather
    if WORKLOAD.startswith("claycat"):       # plasticine_0300 weights
        GOLD = os.path.join(ROOT, "tests", "golden", "material_plasticine.npz")
    if args.impl == "reference":
        args.steps = min(args.steps, 5)         # each step is a ~10-20 s CPU sample
        args.warmup = min(args.warmup, 1)
        run_reference(args)
    elif args.slab:
        run_slab(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
