"""`np.linspace(0, rmax, nbins+1)` -- the reference's RDF bin edges (amep/spatialcor.py:563) -- is
`k * (rmax / nbins)` with the last entry set to `rmax`, in the arithmetic type of `rmax`
(float32 for the default rmax of a float32 box, float64 for a Python float).  DESIGN.md section 9
item 1 builds on this: a kernel can recompute `edges[k]` with one multiply instead of reading a
threshold table from shared memory.  The identity is pinned here for the NumPy version the
golden vectors were made with (tests/golden/VERSIONS.txt)."""
import numpy as np


def test_linspace_is_k_times_step_bit_for_bit():
    rng = np.random.default_rng(2024)
    for _ in range(2000):
        nbins = int(rng.integers(1, 4000))
        rmax = float(rng.uniform(1e-2, 1e4))
        r32 = np.float32(rmax)
        e32 = np.linspace(0.0, r32, nbins + 1)
        assert e32.dtype == np.float32
        mine = np.arange(nbins + 1, dtype=np.float32) * np.float32(r32 / np.float32(nbins))
        mine[-1] = r32
        np.testing.assert_array_equal(e32, mine.astype(np.float32))
        e64 = np.linspace(0.0, rmax, nbins + 1)
        mine64 = np.arange(nbins + 1, dtype=np.float64) * (rmax / nbins)
        mine64[-1] = rmax
        np.testing.assert_array_equal(e64, mine64)
