"""Golden vectors of ``amep.spatialcor.sfiso(mode='std')`` that pin the reference's float32
*summation order*, produced by the UNMODIFIED reference (AMEP 1.3.0).  Run in the build container
only (needs /root/reference):

    python oracle/make_golden_sfiso_std.py

The Debye worker (amep/spatialcor.py:1363-1399) sums its float32 terms in one of two ways: a 2D array
of terms summed over its rows (row after row) while ``4 * pairs * nq <= 1e9`` bytes, else one
``np.sum`` per q (NumPy's pairwise summation).  The 300-particle cases of make_golden.py only reach the
first; these reach the second (one chunk of 2600^2 pairs) and the first with several chunks.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")
sys.path.insert(0, HERE)


def main():
    from refshim import load_reference
    amep = load_reference()
    rng = np.random.default_rng(811)
    n = 2600
    # 2D, one chunk (the default chunksize exceeds N): 6.76e6 pairs x 39 q values -> pairwise sums
    box2 = np.array([[0, 50], [0, 50], [-0.5, 0.5]], dtype=np.float32)
    c2 = np.zeros((n, 3), dtype=np.float32)
    c2[:, :2] = rng.uniform(0, 50, (n, 2)).astype(np.float32)
    c2[11] = c2[4]     # a coincident pair
    s, q = amep.spatialcor.sfiso(c2, box2, n, qmax=5.0, twod=True, mode="std")
    np.savez_compressed(os.path.join(GOLDEN, "sfiso_std_2d_pairwise.npz"), coords=c2, box=box2, S=s, q=q, qmax=5.0, twod=True)
    print("2d pairwise", s.dtype, len(q), 4 * (n * n - n - 2) * len(q) / 1e9)
    # the same frame in chunks of 1000 (the default of evaluate.SFiso): rows branch, three chunks
    s, q = amep.spatialcor.sfiso(c2, box2, n, qmax=5.0, twod=True, mode="std", chunksize=1000)
    np.savez_compressed(os.path.join(GOLDEN, "sfiso_std_2d_chunks1000.npz"), coords=c2, box=box2, S=s, q=q, qmax=5.0,
                        twod=True, chunksize=1000)
    print("2d chunks", s.dtype, len(q), 4 * (n * 1000) * len(q) / 1e9)
    # 3D, one chunk, pairwise sums of np.sinc terms
    box3 = np.array([[0, 13], [0, 13], [0, 13]], dtype=np.float32)
    c3 = rng.uniform(0, 13, (n, 3)).astype(np.float32)
    s, q = amep.spatialcor.sfiso(c3, box3, n, qmax=20.0, twod=False, mode="std")
    np.savez_compressed(os.path.join(GOLDEN, "sfiso_std_3d_pairwise.npz"), coords=c3, box=box3, S=s, q=q, qmax=20.0, twod=False)
    print("3d pairwise", s.dtype, len(q), 4 * (n * n - n) * len(q) / 1e9)


if __name__ == "__main__":
    main()
