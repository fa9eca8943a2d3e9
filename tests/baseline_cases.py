"""Seeded synthetic frames of the BASELINE configs:
``rng = np.random.default_rng([cfg_id, frame_idx])``, positions drawn in float64 and stored as
float32 (trajectory regime); a float64 copy serves the float64-regime tests.

cfg ids follow BASELINE.md (1-based): cfg1 = BASELINE.json configs[0], ..., cfg5 = configs[4].
"""
import numpy as np

RHO_2D_ABP = 0.6366          # packing fraction 0.5 of unit discs


def box2d(L, dtype=np.float32):
    return np.array([[0, L], [0, L], [-0.5, 0.5]], dtype=dtype)


def box3d(L, dtype=np.float32):
    return np.array([[0, L], [0, L], [0, L]], dtype=dtype)


def cfg1_frame(frame, n=4096):
    """2D uniform, rho = 1, L = 64 (n = 4096)."""
    L = float(np.sqrt(n))
    rng = np.random.default_rng([1, frame])
    c = np.zeros((n, 3))
    c[:, :2] = rng.uniform(0, L, (n, 2))
    return c, L


def cfg2_frame(frame, n=1_000_000, clustered=False, nblobs=100, sigma=20.0):
    """2D ABP-like: uniform at rho = 0.6366 (L = 1253.3 for 1M); ``clustered``: half of the
    particles in Gaussian blobs (sigma_blob = 20) wrapped into the box, half uniform."""
    L = float(np.sqrt(n / RHO_2D_ABP))
    rng = np.random.default_rng([2, frame])
    c = np.zeros((n, 3))
    if clustered:
        half = n // 2
        centres = rng.uniform(0, L, (nblobs, 2))
        blob = centres[rng.integers(0, nblobs, half)] + rng.normal(0, sigma, (half, 2))
        c[:half, :2] = np.mod(blob, L)
        c[half:, :2] = rng.uniform(0, L, (n - half, 2))
    else:
        c[:, :2] = rng.uniform(0, L, (n, 2))
    return c, L


def cfg3_frame(frame, m=100, rho=0.8, jitter=0.12):
    """3D LJ-like: FCC lattice of m^3 cells at density rho, Gaussian jitter of ``jitter`` lattice
    constants, wrapped into the box.  N = 4 m^3 (m = 100: 4M, L = 171.0)."""
    a = (4.0 / rho) ** (1.0 / 3.0)
    L = m * a
    rng = np.random.default_rng([3, frame])
    base = np.array([[0, 0, 0], [0.5, 0.5, 0], [0.5, 0, 0.5], [0, 0.5, 0.5]])
    ii = np.arange(m, dtype=np.float64)
    cell = np.stack(np.meshgrid(ii, ii, ii, indexing="ij"), -1).reshape(-1, 1, 3)
    pos = ((cell + base[None]) * a).reshape(-1, 3)
    pos += rng.normal(0, jitter * a, pos.shape)
    return np.mod(pos, L), L


def cfg4_frame(frame, n=262_144):
    """2D uniform rho = 1 (L = 512 for 262 144)."""
    L = float(np.sqrt(n))
    rng = np.random.default_rng([4, frame])
    c = np.zeros((n, 3))
    c[:, :2] = rng.uniform(0, L, (n, 2))
    return c, L


def cfg5_frame(frame, n=1 << 24, rho=0.8):
    """3D uniform rho = 0.8 (L = 275.8 for 2^24)."""
    L = float((n / rho) ** (1.0 / 3.0))
    rng = np.random.default_rng([5, frame])
    return rng.uniform(0, L, (n, 3)), L


def corner_sample(coords, L, side, rmax, dim):
    """Queries = the particles inside the corner cube / square [0, side)^dim, and the particles
    whose periodic images can lie within ``rmax`` of one of them (a conservative pre-filter; the
    brute-force oracle over these images gives exactly the counts of the full frame, because every
    dropped image is farther than rmax from every query).  The corner placement makes the
    sample's neighbourhoods cross the periodic boundary in every direction.

    Returns ``(query_index, neighbour_index)`` into ``coords``."""
    c = np.asarray(coords)
    pad = rmax * 1.001 + 1e-3
    q = np.ones(len(c), dtype=bool)
    near = np.ones(len(c), dtype=bool)
    for d in range(dim):
        x = c[:, d]
        q &= (x >= 0) & (x < side)
        near &= (x <= side + pad) | (x >= L - pad)
    return np.flatnonzero(q), np.flatnonzero(near)
