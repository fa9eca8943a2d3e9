"""Quick device-resident timing of amepb_rdf_batch for a few shapes (development aid)."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import amep_b200
from amep_b200 import _core, _lib


def run(name, n, nf, dim, L, rmax, nbins, dtype=torch.float32, ks=(0,), reps=3, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    c = torch.zeros((nf, n, 3), device="cuda", dtype=dtype)
    c[:, :, :dim] = (torch.rand((nf, n, dim), generator=g, device="cuda", dtype=torch.float64) * L).to(dtype)
    lz = [-0.5, 0.5] if dim == 2 else [0, L]
    box = np.array([[0, L], [0, L], lz], dtype=np.float32 if dtype == torch.float32 else np.float64)
    edges = np.linspace(0.0, rmax, nbins + 1)
    ctx = _lib.context(0)
    ctx.set_timing(True)
    for k in ks:
        ctx.set_cells_per_cutoff(k)
        hist, dims = _core.rdf_histograms(c, box, edges)   # warm-up (allocations)
        torch.cuda.synchronize()
        best = 1e30
        all_ms = []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            hist, dims = _core.rdf_histograms(c, box, edges)
            e1.record()
            torch.cuda.synchronize()
            all_ms.append(round(e0.elapsed_time(e1), 3))
            best = min(best, e0.elapsed_time(e1))
        info = ctx.rdf_info()
        pairs = int(hist.sum().item())
        print(json.dumps(dict(case=name, k=k, k_used=info["cells_per_cutoff"], n=n, frames=nf, ms=round(best, 3),
                              ms_per_frame=round(best / nf, 4), all_ms=all_ms, pair_ms=round(info["pair_kernel_ms"], 3),
                              build_ms=round(info["build_ms"], 3), in_range_pairs=pairs,
                              pair_evals_per_s=pairs / best * 1e3, candidates=info["candidate_pairs"],
                              cand_per_s=info["candidate_pairs"] / best * 1e3,
                              efficiency=round(pairs / max(1, info["candidate_pairs"]), 3),
                              images=info["images"], cells=info["cells"], tasks=info["block_tasks"],
                              launches=info["kernel_launches"], sub_batches=info["sub_batches"])), flush=True)
    ctx.set_cells_per_cutoff(0)


if __name__ == "__main__":
    which = sys.argv[1:] or ["cfg1", "cfg2", "cfg5"]
    if "cfg1" in which:
        run("cfg1_2d_4096x100_rmaxL2", 4096, 100, 2, 64.0, 32.0, 500, ks=(0, 2, 3, 4, 6))
    if "cfg2" in which:
        L = float(np.sqrt(1_000_000 / 0.6366))
        run("cfg2_2d_1M_rmax10", 1_000_000, 16, 2, L, 10.0, 500, ks=(0, 1, 2, 3, 4))
        run("cfg2_2d_1M_rmax10_f64", 1_000_000, 8, 2, L, 10.0, 500, dtype=torch.float64, ks=(0,))
    if "cfg2fast" in which:
        L = float(np.sqrt(1_000_000 / 0.6366))
        run("cfg2_2d_1M_rmax10 KX=%s KY=%s" % (os.environ.get("AMEPB_KX", "-"), os.environ.get("AMEPB_KY", "-")),
            1_000_000, 32, 2, L, 10.0, 500, ks=(0,))
    if "cfg2clustered" in which:
        n = 1_000_000
        L = float(np.sqrt(n / 0.6366))
        g = torch.Generator(device="cuda").manual_seed(3)
        nf = 8
        c = torch.zeros((nf, n, 3), device="cuda", dtype=torch.float32)
        centres = torch.rand((nf, 100, 2), generator=g, device="cuda", dtype=torch.float64) * L
        which_blob = torch.randint(0, 100, (nf, n // 2), generator=g, device="cuda")
        blob = torch.gather(centres, 1, which_blob[..., None].expand(-1, -1, 2)) + torch.randn((nf, n // 2, 2), generator=g, device="cuda", dtype=torch.float64) * 20.0
        c[:, : n // 2, :2] = torch.remainder(blob, L).float()
        c[:, n // 2:, :2] = (torch.rand((nf, n - n // 2, 2), generator=g, device="cuda", dtype=torch.float64) * L).float()
        box = np.array([[0, L], [0, L], [-0.5, 0.5]], dtype=np.float32)
        edges = np.linspace(0.0, 10.0, 501)
        ctx = _lib.context(0)
        ctx.set_timing(True)
        _core.rdf_histograms(c, box, edges)
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            hist, _ = _core.rdf_histograms(c, box, edges)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        info = ctx.rdf_info()
        pairs = int(hist.sum().item())
        print(json.dumps(dict(case="cfg2ii_clustered KX=%s KY=%s" % (os.environ.get("AMEPB_KX", "-"), os.environ.get("AMEPB_KY", "-")),
                              ms_per_frame=round(best / nf, 4), pair_ms=round(info["pair_kernel_ms"] / nf, 4),
                              build_ms=round(info["build_ms"] / nf, 4), pair_evals_per_s=pairs / best * 1e3,
                              candidates=info["candidate_pairs"])), flush=True)
    if "cfg5fast" in which:
        n = 2_000_000
        L = float((n / 0.8) ** (1 / 3))
        run("cfg5like_3d_2M_rmax10", n, 1, 3, L, 10.0, 500, ks=(0,), reps=3)
    if "cfg5" in which:
        n = 2_000_000
        L = float((n / 0.8) ** (1 / 3))
        run("cfg5like_3d_2M_rmax10", n, 1, 3, L, 10.0, 500, ks=(0, 2, 3, 4), reps=2)
