"""Host-vs-kernel time of amepb_rdf_batch on the clustered cfg2(ii) shape (development aid)."""
import os, sys, time, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from amep_b200 import _core, _lib
ctx = _lib.context(0); ctx.set_timing(True)
g = torch.Generator(device="cuda").manual_seed(3)
n = 1_000_000; L = float(np.sqrt(n / 0.6366))
edges = np.linspace(0, 10.0, 501); box = np.array([[0, L], [0, L], [-0.5, 0.5]], dtype=np.float32)
def make(nf, clustered):
    c = torch.zeros((nf, n, 3), device="cuda", dtype=torch.float32)
    if clustered:
        centres = torch.rand((nf, 100, 2), generator=g, device="cuda", dtype=torch.float64) * L
        which = torch.randint(0, 100, (nf, n // 2), generator=g, device="cuda")
        blob = torch.gather(centres, 1, which[..., None].expand(-1, -1, 2)) + torch.randn((nf, n // 2, 2), generator=g, device="cuda", dtype=torch.float64) * 20.0
        c[:, : n // 2, :2] = torch.remainder(blob, L).float()
        c[:, n // 2:, :2] = (torch.rand((nf, n - n // 2, 2), generator=g, device="cuda", dtype=torch.float64) * L).float()
    else:
        c[:, :, :2] = (torch.rand((nf, n, 2), generator=g, device="cuda", dtype=torch.float64) * L).float()
    return c
for clustered in (False, True):
    for nf in (1, 8):
        c = make(nf, clustered)
        for k in (0, 2, 3, 4):
            ctx.set_cells_per_cutoff(k)
            _core.rdf_histograms(c, box, edges); torch.cuda.synchronize()
            t0 = time.perf_counter(); h, _ = _core.rdf_histograms(c, box, edges); torch.cuda.synchronize(); wall = (time.perf_counter() - t0) * 1e3
            i = ctx.rdf_info()
            print(json.dumps(dict(clustered=clustered, nf=nf, k=k, k_used=i["cells_per_cutoff"], wall_ms=round(wall, 3), pair_ms=round(i["pair_kernel_ms"], 3), build_ms=round(i["build_ms"], 3),
                                  tasks=i["block_tasks"], cells=i["cells"], cand=i["candidate_pairs"], pairs=int(h.sum().item()))), flush=True)
        ctx.set_cells_per_cutoff(0)
        del c
