"""Debug aid: symmetric vs ordered vs oracle on the 2d_f32_smallr shape for several seeds."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import c_oracle as corc
from amep_b200 import _lib, spatialcor as sc
box = np.array([[0, 80], [0, 80], [-0.5, 0.5]], dtype=np.float32)
ctx = _lib.context(0)
for seed in range(int(sys.argv[1]) if len(sys.argv) > 1 else 6):
    rng = np.random.default_rng(seed)
    c = np.zeros((6000, 3), dtype=np.float32)
    c[:, :2] = rng.uniform(0, 80, (6000, 2)).astype(np.float32)
    want, _ = corc.rdf_counts(c, box, rmax=5.0, nbins=300)
    res = {}
    for k in (0, 1, 2, 3):
        ctx.set_cells_per_cutoff(k)
        for mode in (True, False):
            ctx.set_symmetric(mode)
            _, _, counts = sc.rdf_frames(c, box, rmax=5.0, nbins=300, return_counts=True)
            d = counts[0].astype(np.int64) - want
            res[(k, mode)] = (int(np.abs(d).sum()), np.nonzero(d)[0][:6].tolist(), d[np.nonzero(d)[0][:6]].tolist())
    ctx.set_symmetric(True); ctx.set_cells_per_cutoff(0)
    print(seed, res, flush=True)
