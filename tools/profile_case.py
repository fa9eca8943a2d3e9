"""One warm-up call and one measured call of a single RDF case (for ncu)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from amep_b200 import _core, _lib

case = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
k = int(sys.argv[2]) if len(sys.argv) > 2 else 0
g = torch.Generator(device="cuda").manual_seed(0)
if case == "cfg2":
    n, nf, dim, rmax = 1_000_000, 4, 2, 10.0
    L = float(np.sqrt(n / 0.6366))
elif case == "cfg1":
    n, nf, dim, rmax, L = 4096, 100, 2, 32.0, 64.0
elif case == "3d":
    n, nf, dim, rmax = 1_000_000, 1, 3, 10.0
    L = float((n / 0.8) ** (1 / 3))
else:
    raise SystemExit("unknown case")
c = torch.zeros((nf, n, 3), device="cuda", dtype=torch.float32)
c[:, :, :dim] = (torch.rand((nf, n, dim), generator=g, device="cuda", dtype=torch.float64) * L).float()
box = np.array([[0, L], [0, L], [-0.5, 0.5] if dim == 2 else [0, L]], dtype=np.float32)
edges = np.linspace(0.0, rmax, 501)
ctx = _lib.context(0)
ctx.set_cells_per_cutoff(k)
for _ in range(2):
    hist, dims = _core.rdf_histograms(c, box, edges)
    torch.cuda.synchronize()
print(case, "pairs", int(hist.sum().item()), ctx.rdf_info())
