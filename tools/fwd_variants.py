"""Accuracy / speed of the constitutive forward kernel for different (bf16 splits, pipelines).
  python tools/fwd_variants.py --build      (CPU box: builds tools/bin/libneuma_fwd_<ns>_<np>.so)
  python tools/fwd_variants.py <ns> <np>    (GPU box)"""
import glob, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VARS = [(3, 3), (3, 2), (2, 3), (2, 4)]
def so(ns, np_): return os.path.join(ROOT, "tools", "bin", f"libneuma_fwd_{ns}_{np_}.so")
if "--build" in sys.argv:
    src = sorted(glob.glob(os.path.join(ROOT, "neuma_b200", "csrc", "*.cu")))
    for ns, np_ in VARS:
        subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
                        "-shared", "--expt-relaxed-constexpr", f"-DNEUMA_NS_FWD={ns}", f"-DNEUMA_NPIPE_FWD={np_}",
                        "-o", so(ns, np_)] + src, check=True)
        print(so(ns, np_))
    sys.exit(0)
ns, np_ = int(sys.argv[1]), int(sys.argv[2])
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from neuma_b200 import _lib
_lib.LIB_PATH = so(ns, np_)
from neuma_b200 import ops
from oracle import constitutive_oracle as CO
dev = torch.device("cuda:0")
z = np.load(os.path.join(ROOT, "tests", "golden", "material_jelly.npz"))
res = {"ns": ns, "npipe": np_}
for tag, kind in (("elastic", 0), ("plastic", 1)):
    W = [torch.from_numpy(z[f"param_{tag}.{l}.fc.weight"]).double() for l in ("layers.0", "layers.1", "final_layer")]
    torch.manual_seed(0)
    F = torch.eye(3, dtype=torch.float64) + 0.08 * torch.randn(20000, 3, 3, dtype=torch.float64)
    alpha = float(z["alpha"])
    ref = CO.elasticity(F, W) if kind == 0 else CO.plasticity(F, W, alpha)
    out = ops.constitutive_fwd(kind, F.float().to(dev), *[w.float().to(dev) for w in W], alpha=alpha if kind else 0.0)
    e = (out.cpu().double() - ref)
    if kind == 1:   # compare the increment F' - F, the part the MLP produces
        res["plastic_rel_err_of_increment"] = float(e.abs().max() / (ref - F).abs().max())
    res[f"{tag}_max_rel_err"] = float(e.abs().max() / ref.abs().max())
n = 100000
We = [torch.from_numpy(z[f"param_elastic.{l}.fc.weight"]).float().to(dev) for l in ("layers.0", "layers.1", "final_layer")]
st = ops.constitutive_stage(*We)
F = (torch.eye(3) + 0.05 * torch.randn(n, 3, 3)).to(dev)
out = torch.empty_like(F)
for _ in range(5): ops.constitutive_fwd_staged(0, F, st, 0.0, out)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(50): ops.constitutive_fwd_staged(0, F, st, 0.0, out)
b.record(); b.synchronize()
res["fwd_us_back_to_back"] = a.elapsed_time(b) * 1e3 / 50
print(res)
