"""Time of sc.pcf_angle (rmax = 10 and the default rmax = L/2) and sc.spatialcor on one 2D frame of N
particles (density 1, float32), the workload of the `sq.pcf_angle_*` / `sq.spatialcor_velocity` bench
entries; also the command ncu captures.  usage: profile_paircor.py [N] [pcfa|scor|all]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from amep_b200 import spatialcor as sc
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
which = sys.argv[2] if len(sys.argv) > 2 else "all"
dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev).manual_seed(7)
L = float(np.sqrt(n))
c = torch.zeros((n, 3), device=dev, dtype=torch.float32)
c[:, :2] = torch.rand((n, 2), generator=gen, device=dev, dtype=torch.float32) * L
box = np.array([[0, L], [0, L], [-0.5, 0.5]], dtype=np.float32)
vel = torch.randn((n, 3), generator=gen, device=dev, dtype=torch.float32)


def one(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


if which in ("pcfa", "all"):
    for tag, kw in (("rmax10", dict(rmax=10.0)), ("default_rmax", dict())):
        ms = one(lambda: sc.pcf_angle(c, box, psi=np.array([0.8, 0.3]), **kw))
        print(f"pcf_angle {tag}: {ms:.2f} ms, {n * (n - 1) / ms * 1e3:.3e} pair evaluations/s")
if which in ("scor", "all"):
    ms = one(lambda: sc.spatialcor(c, box, vel))
    print(f"spatialcor velocity: {ms:.2f} ms, {n * n / ms * 1e3:.3e} pair evaluations/s")
