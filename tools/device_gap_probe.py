"""Device-resident call: wall time vs the sum of the library's own kernel timings (development aid)."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from amep_b200 import _core, _lib
n, nf = 1_000_000, int(sys.argv[1]) if len(sys.argv) > 1 else 512
L = float(np.sqrt(n / 0.6366))
g = torch.Generator(device="cuda").manual_seed(0)
c = torch.zeros((nf, n, 3), device="cuda", dtype=torch.float32)
for f0 in range(0, nf, 64):
    c[f0:f0 + 64, :, :2] = (torch.rand((min(64, nf - f0), n, 2), generator=g, device="cuda", dtype=torch.float64) * L).float()
box = np.array([[0, L], [0, L], [-0.5, 0.5]], dtype=np.float32)
edges = np.linspace(0.0, 10.0, 501)
ctx = _lib.context(0); ctx.set_timing(True)
out = torch.zeros((nf, 500), dtype=torch.int64, device="cuda")
for rep in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); _core.rdf_histograms(c, box, edges, out=out); e1.record(); torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) * 1e3
    i = ctx.rdf_info()
    print(f"rep {rep}: wall {wall:.2f} ms, cuda events {e0.elapsed_time(e1):.2f} ms, pair {i['pair_kernel_ms']:.2f} + build {i['build_ms']:.2f} = {i['pair_kernel_ms'] + i['build_ms']:.2f} ms, sub_batches {i['sub_batches']}", flush=True)
