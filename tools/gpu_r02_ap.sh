#!/bin/bash
# round 2, GPU call AP: partitioned build forced / not forced for small batches (1, 2, 4 frames of 1M particles)
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
python - <<'PY'
import os, sys, numpy as np, torch
sys.path.insert(0, os.getcwd())
from amep_b200 import _core, _lib
n = 1_000_000
L = float(np.sqrt(n / 0.6366))
g = torch.Generator(device="cuda").manual_seed(0)
box = np.array([[0, L], [0, L], [-0.5, 0.5]], dtype=np.float32)
edges = np.linspace(0.0, 10.0, 501)
ctx = _lib.context(0); ctx.set_timing(True)
for nf in (1, 2, 4):
    c = torch.zeros((nf, n, 3), device="cuda", dtype=torch.float32)
    c[:, :, :2] = (torch.rand((nf, n, 2), generator=g, device="cuda", dtype=torch.float64) * L).float()
    for mode in ("auto", "forced", "placement"):
        for v in ("AMEPB_FORCE_PARTITIONED_BUILD", "AMEPB_NO_PARTITIONED_BUILD"):
            os.environ.pop(v, None)
        if mode == "forced": os.environ["AMEPB_FORCE_PARTITIONED_BUILD"] = "1"
        if mode == "placement": os.environ["AMEPB_NO_PARTITIONED_BUILD"] = "1"
        best = 1e9
        for _ in range(5):
            _core.rdf_histograms(c, box, edges); torch.cuda.synchronize()
            best = min(best, ctx.rdf_info()["build_ms"] / nf)
        print(f"frames {nf} {mode}: build {best * 1e3:.1f} us per frame")
PY
