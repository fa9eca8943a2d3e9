"""CUDA-event time of the cell build of configs[1] frames (stats + count + scan + fill), 64 frames per call."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from amep_b200 import _core, _lib
n, nf = 1_000_000, 64
L = float(np.sqrt(n / 0.6366))
g = torch.Generator(device="cuda").manual_seed(0)
c = torch.zeros((nf, n, 3), device="cuda", dtype=torch.float32)
c[:, :, :2] = (torch.rand((nf, n, 2), generator=g, device="cuda", dtype=torch.float64) * L).float()
box = np.array([[0, L], [0, L], [-0.5, 0.5]], dtype=np.float32)
edges = np.linspace(0.0, 10.0, 501)
ctx = _lib.context(0); ctx.set_timing(True)
best = 1e9
for _ in range(4):
    try:
        _core.rdf_histograms(c, box, edges)
    except Exception as exc:
        print("call failed:", exc)
    torch.cuda.synchronize()
    i = ctx.rdf_info()
    best = min(best, i["build_ms"] / nf)
print(f"build {best * 1e3:.2f} us per frame; pair {i['pair_kernel_ms'] / nf * 1e3:.1f} us per frame")
