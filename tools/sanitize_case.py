"""Small end-to-end exercise of every kernel family (for compute-sanitizer)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from amep_b200 import spatialcor as sc
rng = np.random.default_rng(0)
box2 = np.array([[0, 30], [0, 30], [-0.5, 0.5]], dtype=np.float32)
c2 = np.zeros((3, 1500, 3), dtype=np.float32); c2[:, :, :2] = rng.uniform(0, 30, (3, 1500, 2))
g, r, h = sc.rdf_frames(torch.from_numpy(c2).cuda(), box2, rmax=4.0, nbins=100, return_counts=True)      # SYM f32
g, r, h2 = sc.rdf_frames(c2.astype(np.float64), box2.astype(np.float64), rmax=4.0, nbins=100, return_counts=True)  # f64, host space
box3 = np.array([[0, 12], [0, 12], [0, 12]], dtype=np.float32)
c3 = rng.uniform(0, 12, (2, 1200, 3)).astype(np.float32); o3 = rng.uniform(0, 12, (2, 500, 3)).astype(np.float32)
sc.rdf_frames(c3, box3, o3, rmax=3.0, nbins=64)            # cross, 3D
sc.rdf_frames(c3, box3, rmax=3.0, nbins=5000)              # global-histogram path
sc.rdf_frames(c3.astype(np.float64), box3, pbc=False, rmax=3.0, nbins=64)
sc.sfiso_frames(c3, box3, None, qmax=8.0, twod=False, mode="fast", num=8)
sc.sfiso_frames(c2[:, :300], box2, None, qmax=6.0, twod=True, mode="std")
sc.sf2d_frames(c2[:, :600], box2, qmax=3.0, chunksize=97)
sc.sf2d_frames(c2[:, :600], box2, qmax=3.0, accum="f64")
torch.cuda.synchronize()
print("sanitize case done", int(h.sum()), int(h2.sum()))
