"""Forty back-to-back host-buffer calls with wall-clock stamps: are the slow ones periodic?"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from amep_b200 import _core, _lib
n, nf = 1_000_000, 256
L = float(np.sqrt(n / 0.6366))
host = torch.empty((nf, n, 3), dtype=torch.float32).pin_memory()
g = torch.Generator(device="cuda").manual_seed(0)
for f0 in range(0, nf, 64):
    t = torch.zeros((64, n, 3), device="cuda", dtype=torch.float32)
    t[:, :, :2] = (torch.rand((64, n, 2), generator=g, device="cuda", dtype=torch.float64) * L).float()
    host[f0:f0 + 64].copy_(t)
del t
hn = host.numpy()
box = np.array([[0, L], [0, L], [-0.5, 0.5]], dtype=np.float32)
edges = np.linspace(0.0, 10.0, 501)
out = torch.empty((nf, 500), dtype=torch.int64).pin_memory(); on = out.numpy()
ctx = _lib.context(0); ctx.set_timing(True)
_core.rdf_histograms(hn, box, edges, out=on)
t_start = time.perf_counter()
rows = []
for rep in range(40):
    t0 = time.perf_counter(); _core.rdf_histograms(hn, box, edges, out=on); t1 = time.perf_counter()
    i = ctx.rdf_info()
    rows.append((round((t0 - t_start) * 1e3), round((t1 - t0) * 1e3, 1), round(i["pair_kernel_ms"] + i["build_ms"], 1)))
print("start_ms, wall_ms, kernel_ms:", rows)
