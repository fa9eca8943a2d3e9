"""Prototype of the table-free RDF binning proposed in DESIGN.md section 9 item 1, checked against
np.histogram's own semantics on the CPU (float32 regime: float32 distances, float32 edges from
`np.linspace(0, rmax32, nbins+1)`).

    g    = rint(s * inv_w)                      (the kernel's magic-number rounding), 0 <= g <= nbins
    e_g  = rmax if g == nbins else fl32(g * step)          (edges[g], see test_linspace_edges_cpu.py)
    bin  = g - 1 + (s >= e_g)                              (bin nbins: s == rmax goes to the last bin)

The true bin is g - 1 or g because s * inv_w is within 1e-3 of s / step for nbins <= 20000.  Exercised
on random distances and on every edge value with its float32 neighbours."""
import numpy as np

f32 = np.float32


def _reference_bins(s, edges):
    """np.histogram with explicit edges: searchsorted(..., 'right') - 1, closed last bin, -1 = dropped."""
    e = edges.astype(np.float64)
    v = s.astype(np.float64)
    b = np.searchsorted(e, v, side="right") - 1
    b[v == e[-1]] = len(e) - 2
    b[(v < e[0]) | (v > e[-1])] = -1
    return b


def _arith_bins(s, rmax32, nbins):
    step = f32(rmax32 / f32(nbins))
    inv_w = f32(f32(1.0) / step)
    g = np.rint((s * inv_w).astype(f32)).astype(np.int64)
    g = np.clip(g, 0, nbins)
    e_g = (g.astype(f32) * step).astype(f32)
    e_g[g == nbins] = rmax32
    b = g - 1 + (s >= e_g)
    out = np.where(b >= nbins, np.where(s == rmax32, nbins - 1, -1), b)
    return out


def test_arithmetic_thresholds_bin_like_np_histogram():
    rng = np.random.default_rng(5150)
    for _ in range(300):
        nbins = int(rng.integers(1, 20001))
        rmax32 = f32(rng.uniform(0.5, 3000.0))
        edges = np.linspace(0.0, rmax32, nbins + 1)
        assert edges.dtype == np.float32
        s = (rng.uniform(0, 1.02, 20000) * float(rmax32)).astype(f32)
        pick = edges[rng.integers(0, nbins + 1, 3000)]
        probes = np.concatenate([s, pick, np.nextafter(pick, f32(np.inf)), np.nextafter(pick, f32(-np.inf)),
                                 np.array([0.0, rmax32], dtype=f32)])
        probes = probes[probes >= 0]
        np.testing.assert_array_equal(_arith_bins(probes, rmax32, nbins), _reference_bins(probes, edges))
