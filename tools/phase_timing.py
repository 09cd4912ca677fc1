"""Per-phase clock64() breakdown of one pipeline of constitutive_bwd_kernel (CTA 0, thread 0).
Build (CPU box):  python tools/phase_timing.py --build      -> tools/bin/libneuma_timing.so
Run (GPU box):    python tools/phase_timing.py"""
import ctypes, glob, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "tools", "bin", "libneuma_timing.so")
if "--build" in sys.argv:
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    src = sorted(glob.glob(os.path.join(ROOT, "neuma_b200", "csrc", "*.cu")))
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-shared",
           "--expt-relaxed-constexpr", "-DNEUMA_PHASE_TIMING", "-o", SO] + src
    subprocess.run(cmd, check=True)
    print(SO)
    sys.exit(0)
sys.path.insert(0, ROOT)
import numpy as np, torch
from neuma_b200 import _lib
_lib.LIB_PATH = SO
from neuma_b200 import ops
dev = torch.device("cuda:0")
n = int(os.environ.get("N", "100000"))
z = np.load(os.path.join(ROOT, "tests", "golden", "material_jelly.npz"))
W = [torch.from_numpy(z[f"param_elastic.{l}.fc.weight"]).float().to(dev) for l in ("layers.0", "layers.1", "final_layer")]
torch.manual_seed(0)
F = (torch.eye(3) + 0.05 * torch.randn(n, 3, 3)).to(dev)
g = torch.randn(n, 3, 3, device=dev)
gF = torch.empty_like(F)
gW = [torch.zeros_like(w) for w in W]
st = ops.constitutive_stage(*W)
lib = _lib.load()
lib.neuma_debug_phase_cycles.argtypes = [ctypes.c_void_p, ctypes.c_int]
names = ["kinematics", "feat+publish0", "wait0(h1)", "gelu1+pub", "wait1(h2)", "gelu2+pub", "wait2(o)", "head+pub", "wait3(da2)",
         "grad2(+wdone)+pub", "wait4(da1)", "grad1(+wdone)+pub", "wait5(dfeat)", "tail(dF)", "final wdone", "top wdone"]
for wg in (True, False):
    for _ in range(2):
        ops.constitutive_bwd_staged(0, F, st, 0.0, g, gF, *(gW if wg else (None, None, None)))
    lib.neuma_debug_phase_cycles(None, 1)
    R = 1
    for _ in range(R):
        ops.constitutive_bwd_staged(0, F, st, 0.0, g, gF, *(gW if wg else (None, None, None)))
    out = (ctypes.c_longlong * 64)()
    lib.neuma_debug_phase_cycles(out, 1)
    tot = sum(out[i] for i in range(16))
    print(f"--- WGRAD={wg}: total {tot / R:.0f} cycles per launch for pipeline 0 of CTA 0")
    for i, nm in enumerate(names):
        print(f"   {nm:22s} {out[i] / R:9.0f}  {out[i] / tot * 100:5.1f}%")
    print(f"   globaltimer: first CTA start -> last CTA end {(out[25] - out[24]) / 1e3:.2f} us; CTA0 start->end "
          f"{(out[28] - out[24]) / 1e3:.2f} us, {out[27] - out[26]} cycles => {(out[27] - out[26]) / max(out[28] - out[24], 1):.3f} GHz")
    print("   MMA warp (CTA 0): issue cycles per job type (sum over jobs) " + ", ".join(f"J{j}={out[32 + j] / R:.0f}" for j in range(6))
          + f"; jobs={out[40] / R:.0f}; warp lifetime {out[41] / R:.0f} cycles")
    for i, nm in zip(range(16, 21), ["setup", "main loop", "final sync wait", "GW readout+atomics", "sync+dealloc"]):
        print(f"   kernel: {nm:22s} {out[i] / R:9.0f}")
