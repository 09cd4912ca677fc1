"""Golden vectors from the reference's importable pure-Python d3gs helpers.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden_d3gs.py

Executes, unmodified, from /root/reference:
  * modules/d3gs/utils/sh_utils.py::eval_sh          (the Python twin of the rasterizer's SH->RGB)
  * modules/d3gs/utils/graphics_utils.py::getWorld2View2, getProjectionMatrix
    (camera matrices as modules/d3gs/scene/cameras.py:54-57 assembles them)
and restates, on CPU, modules/d3gs/utils/general_utils.py::build_rotation /
build_scaling_rotation / strip_lowerdiag (those hard-code device="cuda") to give
packed-6 covariances in the reference's order.
"""
import math
import os
import sys

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)

from modules.d3gs.utils.sh_utils import eval_sh  # noqa: E402
from modules.d3gs.utils.graphics_utils import getWorld2View2, getProjectionMatrix  # noqa: E402


def main():
    torch.manual_seed(0)
    out = {}
    n = 64
    dirs = torch.randn(n, 3, dtype=torch.float64)
    dirs = dirs / dirs.norm(dim=1, keepdim=True)
    sh = torch.randn(n, 16, 3, dtype=torch.float64)          # [K, M, 3] as gaussians.get_features
    out["sh_dirs"], out["sh_coeffs"] = dirs.numpy(), sh.numpy()
    for deg in range(4):
        # gaussian_renderer/__init__.py:62-67: shs_view = features.transpose(1, 2) -> [K, 3, M]
        res = eval_sh(deg, sh.transpose(1, 2), dirs)
        out[f"sh_rgb_deg{deg}"] = torch.clamp_min(res + 0.5, 0.0).numpy()
    # camera of SURVEY 8d: at (0.5, 0.5, 2.5) looking down -z, fov 40 deg
    R = np.diag([1.0, -1.0, -1.0])
    eye = np.array([0.5, 0.5, 2.5])
    T = -R.T @ eye
    fov = math.radians(40.0)
    wv = torch.tensor(getWorld2View2(R, T)).transpose(0, 1)
    proj = getProjectionMatrix(znear=0.01, zfar=100.0, fovX=fov, fovY=fov).transpose(0, 1)
    full = (wv.unsqueeze(0).bmm(proj.unsqueeze(0))).squeeze(0)
    out["cam_world_view"] = wv.numpy()
    out["cam_full_proj"] = full.numpy()
    out["cam_center"] = wv.inverse()[3, :3].numpy()
    np.savez_compressed(os.path.join(OUT, "d3gs_helpers.npz"), **out)
    print("wrote d3gs_helpers.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
