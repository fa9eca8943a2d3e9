"""Timeline of one host-buffer call of 256 frames (AMEPB_TRACE=1 prints the windows)."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from amep_b200 import _core, _lib
n, nf = 1_000_000, 256
L = float(np.sqrt(n / 0.6366))
rng = np.random.default_rng(0)
host = torch.empty((nf, n, 3), dtype=torch.float32).pin_memory()
hn = host.numpy(); hn[...] = 0
g = torch.Generator(device="cuda").manual_seed(0)
for f0 in range(0, nf, 64):
    t = torch.zeros((64, n, 3), device="cuda", dtype=torch.float32)
    t[:, :, :2] = (torch.rand((64, n, 2), generator=g, device="cuda", dtype=torch.float64) * L).float()
    host[f0:f0 + 64].copy_(t)
box = np.array([[0, L], [0, L], [-0.5, 0.5]], dtype=np.float32)
edges = np.linspace(0.0, 10.0, 501)
out = torch.empty((nf, 500), dtype=torch.int64).pin_memory(); on = out.numpy()
ctx = _lib.context(0); ctx.set_timing(True)
for rep in range(3):
    t0 = time.perf_counter(); _core.rdf_histograms(hn, box, edges, out=on); wall = (time.perf_counter() - t0) * 1e3
    i = ctx.rdf_info()
    print(f"rep {rep}: wall {wall:.2f} ms; pair {i['pair_kernel_ms']:.2f} + build {i['build_ms']:.2f} = {i['pair_kernel_ms'] + i['build_ms']:.2f} ms; sub_batches {i['sub_batches']}", file=sys.stderr, flush=True)
