"""sfiso('fast') with the reference's defaults (twod=True, num=8, qmax=20) on configs[1]-sized 2D frames."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from amep_b200 import spatialcor as sc, _lib
ctx = _lib.context(0); ctx.set_timing(True)
g = torch.Generator(device="cuda").manual_seed(1)
for n, nf in ((1_000_000, 2), (262_144, 4)):
    L = float(np.sqrt(n / 0.6366))
    c = torch.zeros((nf, n, 3), device="cuda", dtype=torch.float32)
    c[:, :, :2] = (torch.rand((nf, n, 2), generator=g, device="cuda", dtype=torch.float64) * L).float()
    box = np.array([[0, L], [0, L], [-0.5, 0.5]], dtype=np.float32)
    for _ in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        s, q = sc.sfiso_frames(c, box, None, qmax=20.0, twod=True, mode="fast", num=8)
        torch.cuda.synchronize(); wall = (time.perf_counter() - t0) * 1e3
    nq = q.shape[1]
    print(f"N={n} frames={nf} nq={nq}: wall {wall:.2f} ms, kernel {ctx.last_kernel_ms():.2f} ms, "
          f"terms/s (8 directions) {nf * n * nq * 8 / (wall * 1e-3):.3e}", flush=True)
