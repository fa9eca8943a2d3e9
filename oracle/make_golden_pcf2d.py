"""Golden vectors for spatialcor.pcf2d, produced by the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python oracle/make_golden_pcf2d.py

Stores the seeded inputs and what ``amep.spatialcor.pcf2d`` returned
(amep/spatialcor.py:726-926): g(x, y), X, Y.  The integer pair counts are recovered from g by
undoing the normalisation (hist = g * area * density * N_other, rounded) and stored as well.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")
sys.path.insert(0, HERE)


def _case(amep, name, coords, box, other=None, nxbins=None, nybins=None, rmax=None, psi=None):
    # the reference zeroes the last column of its arguments in place: hand it copies
    g, X, Y = amep.spatialcor.pcf2d(coords.copy(), box, other_coords=None if other is None else other.copy(),
                                    nxbins=nxbins, nybins=nybins, rmax=rmax, psi=psi)
    nx, ny = g.shape
    r = max(box[:, 1] - box[:, 0]) // (2 * np.sqrt(2)) if rmax is None else rmax
    xb, yb = np.linspace(-r, r, nx + 1), np.linspace(-r, r, ny + 1)
    area = (xb[1:] - xb[:-1])[:, None] * (yb[1:] - yb[:-1])
    bl = box[:, 1] - box[:, 0]
    n_other = len(coords) if other is None else len(other)
    counts = np.rint(g * area * (len(coords) / (bl[0] * bl[1])) * n_other).astype(np.int64)
    out = dict(coords=coords, box=box, g=g, X=X, Y=Y, counts=counts,
               nxbins=np.int64(-1 if nxbins is None else nxbins), nybins=np.int64(-1 if nybins is None else nybins),
               rmax=np.float64(np.nan if rmax is None else rmax))
    if other is not None:
        out["other"] = other
    if psi is not None:
        out["psi"] = np.asarray(psi, dtype=np.float64)
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)
    print(f"{name}: g {g.shape}, pairs in range {counts.sum()} of {len(coords) * n_other}, max count {counts.max()}")


def main():
    from refshim import load_reference
    amep = load_reference()
    rng = np.random.default_rng(417)
    box64 = np.array([[0, 30], [0, 30], [-0.5, 0.5]], dtype=np.float64)
    c = np.zeros((700, 3))
    c[:, :2] = rng.uniform(0, 30, (700, 2))
    c[:, 2] = rng.uniform(-0.5, 0.5, 700)      # discarded by the reference
    o = np.zeros((450, 3))
    o[:, :2] = rng.uniform(0, 30, (450, 2))
    # float32, same particles, default rmax = 30 // (2 sqrt 2) = 10, no rotation
    _case(amep, "pcf2d_f32_same", c.astype(np.float32), box64.astype(np.float32), nxbins=40, nybins=40)
    # float64 cross, rotated, different bin numbers
    _case(amep, "pcf2d_f64_cross_rot", c, box64, other=o, nxbins=30, nybins=50, rmax=8.0, psi=np.array([0.3, 0.8]))
    # float32 with a float32 box off the origin (float32 centre), psi[1] < 0
    boxr = np.array([[-10, 25], [-10, 25], [-0.5, 0.5]], dtype=np.float32)
    cr = np.zeros((600, 3), dtype=np.float32)
    cr[:, :2] = rng.uniform(-10, 25, (600, 2)).astype(np.float32)
    _case(amep, "pcf2d_f32_rot_offset", cr, boxr, nxbins=36, nybins=28, rmax=9.5, psi=np.array([0.5, -0.7]))
    # mixed dtypes: float32 coords, float64 other, float64 box
    _case(amep, "pcf2d_mixed_cross", c.astype(np.float32), box64, other=o, nxbins=25, nybins=25, rmax=7.0)
    # lattice: every difference is an integer and sits on a bin edge, the outermost ones on the
    # closed right edge / the open side of the left one
    gx, gy = np.meshgrid(np.arange(16.0), np.arange(16.0))
    lat = np.zeros((256, 3))
    lat[:, 0], lat[:, 1] = gx.ravel(), gy.ravel()
    boxl = np.array([[0, 16], [0, 16], [-0.5, 0.5]], dtype=np.float64)
    _case(amep, "pcf2d_lattice_edges", lat, boxl, nxbins=10, nybins=10, rmax=5.0)
    _case(amep, "pcf2d_lattice_edges_f32", lat.astype(np.float32), boxl.astype(np.float32), nxbins=10, nybins=20, rmax=5.0)


if __name__ == "__main__":
    main()
