"""One pass over the large BASELINE configs (sizes beyond the oracle): cfg3 (3D FCC + jitter, 4M),
cfg5 (3D uniform, 16.7M) for RDF; properties checked: sum of counts vs ideal gas / lattice density,
symmetric vs ordered evaluation on a sub-box, timing."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from amep_b200 import _core, _lib
ctx = _lib.context(0); ctx.set_timing(True)
def timed(c, box, edges, reps=3):
    _core.rdf_histograms(c, box, edges); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); hist, dims = _core.rdf_histograms(c, box, edges); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return hist, best
g = torch.Generator(device="cuda").manual_seed(3)
# cfg2 (ii): 2D "ABP-like" clustered, N = 1M, half of the particles in 100 Gaussian blobs (sigma 20), half uniform
n = 1_000_000; L = float(np.sqrt(n / 0.6366)); nf = 8
edges = np.linspace(0, 10.0, 501); box = np.array([[0, L], [0, L], [-0.5, 0.5]], dtype=np.float32)
c = torch.zeros((nf, n, 3), device="cuda", dtype=torch.float32)
centres = torch.rand((nf, 100, 2), generator=g, device="cuda", dtype=torch.float64) * L
which = torch.randint(0, 100, (nf, n // 2), generator=g, device="cuda")
blob = torch.gather(centres, 1, which[..., None].expand(-1, -1, 2)) + torch.randn((nf, n // 2, 2), generator=g, device="cuda", dtype=torch.float64) * 20.0
c[:, : n // 2, :2] = torch.remainder(blob, L).float()
c[:, n // 2 :, :2] = (torch.rand((nf, n - n // 2, 2), generator=g, device="cuda", dtype=torch.float64) * L).float()
del blob, which, centres
torch.cuda.empty_cache()
hist, ms = timed(c, box, edges)
info = ctx.rdf_info(); tot = int(hist.sum().item())
ctx.set_symmetric(False); h2, ms2 = timed(c[:2], box, edges, reps=1); ctx.set_symmetric(True)
print(json.dumps(dict(case="cfg2ii_2d_clustered_1M_8frames", ms_per_frame=ms / nf, pairs=tot, pairs_per_s=tot / ms * 1e3,
                      pair_ms_per_frame=info["pair_kernel_ms"] / nf, build_ms_per_frame=info["build_ms"] / nf, k=info["cells_per_cutoff"],
                      candidates=info["candidate_pairs"], symmetric_equals_ordered=bool(torch.equal(hist[:2], h2)))), flush=True)
del c
# cfg5: 3D uniform rho = 0.8, N = 2^24
n = 1 << 24; L = float((n / 0.8) ** (1 / 3))
c = torch.rand((1, n, 3), generator=g, device="cuda", dtype=torch.float32) * L
box = np.array([[0, L]] * 3, dtype=np.float32); edges = np.linspace(0, 10.0, 501)
hist, ms = timed(c, box, edges)
tot = int(hist.sum().item()); expect = n * 0.8 * 4 / 3 * np.pi * 1000
info = ctx.rdf_info()
print(json.dumps(dict(case="cfg5_3d_16.7M_1frame", ms=ms, pairs=tot, expected=expect, rel=(tot - expect) / expect,
                      pairs_per_s=tot / ms * 1e3, pair_ms=info["pair_kernel_ms"], build_ms=info["build_ms"], k=info["cells_per_cutoff"],
                      images=info["images"], cells=info["cells"])), flush=True)
del c
# cfg3: FCC lattice at rho = 0.8 with Gaussian jitter 0.12 a, N = 4 * 100^3
m = 100; a = (4 / 0.8) ** (1 / 3); L = m * a
base = torch.tensor([[0, 0, 0], [0.5, 0.5, 0], [0.5, 0, 0.5], [0, 0.5, 0.5]], device="cuda", dtype=torch.float64)
ii = torch.arange(m, device="cuda", dtype=torch.float64)
cell = torch.stack(torch.meshgrid(ii, ii, ii, indexing="ij"), -1).reshape(-1, 1, 3)
pos = ((cell + base[None]) * a).reshape(-1, 3)
pos = pos + torch.randn(pos.shape, generator=g, device="cuda", dtype=torch.float64) * 0.12 * a
pos = torch.remainder(pos, L)
c = pos.float()[None].contiguous(); n = c.shape[1]
box = np.array([[0, L]] * 3, dtype=np.float32)
hist, ms = timed(c, box, edges)
tot = int(hist.sum().item()); expect = n * 0.8 * 4 / 3 * np.pi * 1000
info = ctx.rdf_info()
print(json.dumps(dict(case="cfg3_3d_fcc_4M_1frame", n=n, ms=ms, pairs=tot, expected=expect, rel=(tot - expect) / expect,
                      pairs_per_s=tot / ms * 1e3, pair_ms=info["pair_kernel_ms"], build_ms=info["build_ms"], k=info["cells_per_cutoff"])), flush=True)
# symmetric vs ordered on the same frame must agree bit for bit
ctx.set_symmetric(False); h2, ms2 = timed(c, box, edges, reps=1); ctx.set_symmetric(True)
print(json.dumps(dict(case="cfg3 symmetric == ordered", equal=bool(torch.equal(hist, h2)), ordered_ms=ms2)), flush=True)
# first peak of g(r) of the jittered FCC lattice sits at the nearest-neighbour distance a/sqrt(2)
hh = hist[0].cpu().numpy().astype(float); r = 0.5 * (edges[1:] + edges[:-1])
gr = hh / (4 / 3 * np.pi * (edges[1:] ** 3 - edges[:-1] ** 3)) / 0.8 / n
print(json.dumps(dict(case="cfg3 g(r) first peak", r_peak=float(r[np.argmax(gr[:120])]), nn=a / np.sqrt(2))), flush=True)
