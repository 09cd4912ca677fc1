"""ncu target: the constitutive kernels alone at the bench workload's shape (100k particles)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from neuma_b200 import ops

dev = torch.device("cuda:0")
n = int(os.environ.get("N", "100000"))
z = np.load(os.path.join(ROOT, "tests", "golden", "material_jelly.npz"))
W = [torch.from_numpy(z[f"param_elastic.{l}.fc.weight"]).float().to(dev) for l in ("layers.0", "layers.1", "final_layer")]
torch.manual_seed(0)
F = (torch.eye(3) + 0.05 * torch.randn(n, 3, 3)).to(dev)
g = torch.randn(n, 3, 3, device=dev)
out = torch.empty_like(F)
gF = torch.empty_like(F)
gW = [torch.zeros_like(w) for w in W]
for _ in range(3):
    ops.constitutive_fwd_into(0, F, *W, 0.0, out)
    ops.constitutive_bwd_into(0, F, *W, 0.0, g, gF, *gW)
torch.cuda.synchronize()
print("done")
