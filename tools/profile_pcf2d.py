"""Kernel time of spatialcor.pcf2d counts (amepb_pcf2d_counts) with device-resident coordinates.

    python tools/profile_pcf2d.py N [rmax|default] [nbins] [rot]

Uniform 2D system of density 1 (box side sqrt(N)), float32, same particles."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from amep_b200 import _core

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
L = float(np.sqrt(n))
rmax = sys.argv[2] if len(sys.argv) > 2 else "default"
rmax = L // (2 * np.sqrt(2)) if rmax == "default" else float(rmax)
nb = int(sys.argv[3]) if len(sys.argv) > 3 else 500
rot = len(sys.argv) > 4 and sys.argv[4] == "rot"
rng = np.random.default_rng(5)
c = torch.zeros((n, 3), dtype=torch.float32)
c[:, :2] = torch.from_numpy(rng.uniform(0, L, (n, 2)).astype(np.float32))
c = c.cuda()
edges = np.linspace(-rmax, rmax, nb + 1)
cen = np.array([[L / 2, L / 2, 0.0]])
r = np.array([[1.0, np.cos(-0.3), np.sin(-0.3)]]) if rot else None
for _ in range(2):
    counts = _core.pcf2d_counts(c, None, cen, True, r, edges, edges)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 3
e0.record()
for _ in range(reps):
    counts = _core.pcf2d_counts(c, None, cen, True, r, edges, edges)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
inr = int(counts.sum().item())
print(f"pcf2d N={n} rmax={rmax:g} nbins={nb} rot={rot}: {ms:.3f} ms/frame, pairs {n * (n - 1) / ms * 1e3:.3e}/s, "
      f"in-window {inr} ({inr / ms * 1e3:.3e}/s)")
