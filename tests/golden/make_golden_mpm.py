"""Golden vectors for the MPM substep, batch_svd and deform_cov_by_F, produced by executing the
REFERENCE's own kernel bodies (unmodified source, imported from /root/reference) thread by
thread under tests/golden/warp_shim.py -- warp-lang itself is not installable here.

    python tests/golden/make_golden_mpm.py      (build container only)

Kernels executed: MPMModel.p2g / grid_op_noslip / grid_op_freeslip / g2p
(modules/nclaw/sim/mpm.py:321-498), SVDFunction.batch_svd (modules/nclaw/warp/svd.py:59-96),
deform_cov_by_F (modules/d3gs/utils/simulation_utils.py:25-48)."""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import warp_shim as ws  # noqa: E402

wp = ws.install()
sys.path.insert(0, "/root/reference")
# mpm.py:553-571 declares `lin_vel: np.ndarray = np.zeros(3)` in a @dataclass, which Python >= 3.11
# rejects at import time; MPMInitData is not used here, so neutralise the decorator for the import.
import dataclasses  # noqa: E402
import torch  # noqa: E402,F401  (import before the patch: torch itself uses dataclasses)
import einops  # noqa: E402,F401
_dc = dataclasses.dataclass
dataclasses.dataclass = lambda cls=None, **kw: cls if cls is not None else (lambda c: c)
from modules.nclaw.sim.mpm import MPMModel  # noqa: E402
dataclasses.dataclass = _dc
from modules.nclaw.warp.svd import SVDFunction  # noqa: E402
from modules.d3gs.utils.simulation_utils import deform_cov_by_F  # noqa: E402

f32 = np.float32


def run_case(G, n, bc, dt, seed, gravity=(0.0, -9.8, 0.0), bound=1, eps=6e-7, lo=0.08, hi=0.84):
    rng = np.random.default_rng(seed)
    x = (lo + (hi - lo) * rng.random((n, 3))).astype(f32)
    x[:5] = f32(0.02)                       # x * inv_dx < 0.5: int() truncation, not floor
    v = (np.array([0.0, -0.5, 0.0]) + 2.0 * rng.standard_normal((n, 3))).astype(f32)
    C = (2.0 * rng.standard_normal((n, 3, 3))).astype(f32)
    F = (np.eye(3) + 0.05 * rng.standard_normal((n, 3, 3))).astype(f32)
    stress = (20.0 * rng.standard_normal((n, 3, 3))).astype(f32)
    vol = np.full(n, 0.6 ** 3 / n, f32)
    rho = np.full(n, 1000.0, f32)
    clip = np.full(n, 0.1, f32)
    enabled = (rng.random(n) > 0.1).astype(np.int32)
    const = types.SimpleNamespace(num_grids=G, dt=dt, bound=bound, gravity=ws.vec3(*gravity), dx=1 / G,
                                  inv_dx=float(G), eps=eps)
    statics = types.SimpleNamespace(vol=ws.Arr(vol, "s"), rho=ws.Arr(rho, "s"), clip_bound=ws.Arr(clip, "s"),
                                    enabled=ws.Arr(enabled, "s"))
    curr = types.SimpleNamespace(x=ws.Arr(x, "v"), v=ws.Arr(v, "v"), C=ws.Arr(C, "m"), F=ws.Arr(F, "m"),
                                 stress=ws.Arr(stress, "m"))
    x2, v2 = np.zeros_like(x), np.zeros_like(v)
    C2, F2 = np.zeros_like(C), np.tile(np.eye(3, dtype=f32), (n, 1, 1))
    nxt = types.SimpleNamespace(x=ws.Arr(x2, "v"), v=ws.Arr(v2, "v"), C=ws.Arr(C2, "m"), F=ws.Arr(F2, "m"))
    gmv, gv, gm = np.zeros((G, G, G, 3), f32), np.zeros((G, G, G, 3), f32), np.zeros((G, G, G), f32)
    grid = types.SimpleNamespace(mv=ws.Arr(gmv, "v"), v=ws.Arr(gv, "v"), m=ws.Arr(gm, "s"))
    # MPMModel.forward, mpm.py:290-297 (grid.clear() == the zero arrays above)
    ws.launch(MPMModel.p2g, n, [const, statics, curr, grid])
    grid_op = MPMModel.grid_op_noslip if bc == "noslip" else MPMModel.grid_op_freeslip
    ws.launch(grid_op, [G, G, G], [const, grid])
    ws.launch(MPMModel.g2p, n, [const, statics, curr, nxt, grid])
    return dict(G=G, n=n, bc=bc, dt=dt, gravity=np.array(gravity), bound=bound, eps=eps, x=x, v=v, C=C, F=F,
                stress=stress, vol=vol, rho=rho, clip=clip, enabled=enabled, grid_mv=gmv, grid_m=gm, grid_v=gv,
                x2=x2, v2=v2, C2=C2, F2=F2)


def main():
    out = {}
    cases = [("a", dict(G=12, n=400, bc="noslip", dt=1e-3, seed=0)),
             ("b", dict(G=10, n=300, bc="freeslip", dt=5e-4, seed=1, gravity=(3.0, -9.8, -2.0), bound=2, eps=0.0))]
    for tag, kw in cases:
        r = run_case(**kw)
        for k, v in r.items():
            out[f"{tag}_{k}"] = np.asarray(v)
    # batch_svd (svd.py:59-96) incl. reflections
    rng = np.random.default_rng(7)
    A = rng.standard_normal((64, 3, 3)).astype(f32)
    A[:16] *= np.array([1, 1, -1], f32)
    U, S, Vh = np.zeros_like(A), np.zeros((64, 3), f32), np.zeros_like(A)
    ws.launch(SVDFunction.batch_svd, 64, [ws.Arr(A, "m"), ws.Arr(U, "m"), ws.Arr(S, "v"), ws.Arr(Vh, "m")])
    out.update(svd_A=A, svd_U=U, svd_sigma=S, svd_Vh=Vh)
    # deform_cov_by_F (simulation_utils.py:25-48)
    K = 50
    L = rng.standard_normal((K, 3, 3)).astype(f32)
    Sg = L @ L.transpose(0, 2, 1)
    cov6 = np.stack([Sg[:, 0, 0], Sg[:, 0, 1], Sg[:, 0, 2], Sg[:, 1, 1], Sg[:, 1, 2], Sg[:, 2, 2]], -1).astype(f32)
    Fk = (np.eye(3) + 0.3 * rng.standard_normal((K, 3, 3))).astype(f32)
    flat_in, flat_out = cov6.reshape(-1).copy(), np.zeros(K * 6, f32)
    ws.launch(deform_cov_by_F, K, [ws.Arr(flat_in, "s"), ws.Arr(Fk, "m"), ws.Arr(flat_out, "s")])
    out.update(cov_in=cov6, cov_F=Fk, cov_out=flat_out.reshape(K, 6))
    np.savez_compressed(os.path.join(HERE, "mpm_shim.npz"), **out)
    print("wrote mpm_shim.npz:", sorted(out)[:8], "...")


if __name__ == "__main__":
    main()
