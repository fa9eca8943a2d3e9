"""One warm-up and one measured call of an S(q) case (for ncu)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from amep_b200 import spatialcor as sc, _lib
case = sys.argv[1] if len(sys.argv) > 1 else "sfiso"
g = torch.Generator(device="cuda").manual_seed(1)
ctx = _lib.context(0); ctx.set_timing(True)
if case == "sfiso":
    n = 1_000_000; L = float((n / 0.8) ** (1 / 3))
    c = torch.rand((1, n, 3), generator=g, device="cuda", dtype=torch.float32) * L
    box = np.array([[0, L]] * 3, dtype=np.float32)
    for _ in range(2):
        s, q = sc.sfiso_frames(c, box, None, qmax=20.0, twod=False, mode="fast", num=8)
    print("sfiso nq", q.shape[1], "kernel ms", ctx.last_kernel_ms())
else:
    n, L = 262_144, 512.0
    nf = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    c = torch.zeros((nf, n, 3), device="cuda", dtype=torch.float32)
    c[:, :, :2] = torch.rand((nf, n, 2), generator=g, device="cuda", dtype=torch.float32) * L
    box = np.array([[0, L], [0, L], [-0.5, 0.5]], dtype=np.float32)
    accum = "f64" if case == "sf2d_f64" else "c64"
    for _ in range(2):
        s = sc.sf2d_frames(c, box, qmax=256.5 * 2 * np.pi / 512.0, accum=accum)[0]
    print(case, s.shape, "kernel ms", ctx.last_kernel_ms())
