"""Parity on the BASELINE configs themselves (SURVEY.md 8d "parity criteria"): configs[0] in
full through evaluate.RDF; >= 32 768-query samples of the large 3D frames (configs[2], configs[4])
and of the clustered 2D frame (configs[1](ii)) against the brute-force C oracle, float32 and
float64 regimes; bit-identity statistics of the reference-order sf2d kernel."""
import ctypes
import os

import numpy as np
import pytest

import amep_oracle as orc
import c_oracle as corc
from baseline_cases import (box2d, box3d, cfg1_frame, cfg2_frame, cfg3_frame, cfg4_frame, cfg5_frame,
                            corner_sample)
from conftest import GOLDEN

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def sc():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import amep_b200
    return amep_b200.spatialcor


# ------------------------------------------------------------------------------------------
# configs[0]: amep.evaluate.RDF, 2D uniform 4096 particles, 100 frames, nbins=500, default rmax
# ------------------------------------------------------------------------------------------
def test_cfg1_full_evaluate_rdf(sc):
    """All 100 frames (rmax = L/2: the all-images regime, float32 edges) against the C oracle frame
    by frame, frame 0 against the reference's own output, host and device resident."""
    import amep_b200
    nf = 100
    frames = [cfg1_frame(f) for f in range(nf)]
    L = frames[0][1]
    coords = np.stack([c for c, _ in frames]).astype(np.float32)
    box = box2d(L)
    want_counts, want_g = [], []
    for f in range(nf):
        h, edges = corc.rdf_counts(coords[f], box, nbins=500)
        g, r = orc.rdf_normalise(h, edges, coords[f], box, coords.shape[1])
        want_counts.append(h)
        want_g.append(np.array([g, r]))
    want = np.array(want_g)
    assert edges.dtype == np.float32 and np.sum(want_counts[0]) > 1.3e7
    z = np.load(os.path.join(GOLDEN, "rdf_cfg1_frame0_f32.npz"))     # produced by the unmodified reference
    np.testing.assert_array_equal(z["coords"], coords[0])
    np.testing.assert_array_equal(z["counts"], want_counts[0])
    for data in (coords, torch.from_numpy(coords).cuda()):
        traj = amep_b200.TensorTrajectory(data, box)
        ev = amep_b200.evaluate.RDF(traj, skip=0.0, nav=nf, nbins=500)
        np.testing.assert_array_equal(ev.indices, np.arange(nf))
        np.testing.assert_array_equal(ev.counts, np.array(want_counts))
        np.testing.assert_array_equal(ev.frames, want)
        np.testing.assert_array_equal(ev.avg, np.mean(want, axis=0)[0])
        np.testing.assert_array_equal(ev.frames[0, 0], z["g"])
        np.testing.assert_array_equal(ev.frames[0, 1], z["r"])


# ------------------------------------------------------------------------------------------
# query samples of the large frames against the brute-force oracle
# ------------------------------------------------------------------------------------------
def _oracle_sample_counts(coords, box, L, side, rmax, nbins, dim, qdtype):
    """Brute-force counts (C oracle) of the queries in the corner cube against every periodic image
    of the frame that can be within rmax of one of them."""
    qi, ni = corner_sample(coords, L, side, rmax, dim)
    near = np.ascontiguousarray(coords[ni])
    images = corc.images(near, box)                           # float32 images, window of pbc_points
    pad = rmax * 1.001 + 1e-3
    keep = np.ones(len(images), dtype=bool)
    for d in range(dim):
        keep &= (images[:, d] >= -pad) & (images[:, d] <= side + pad)
    images = np.ascontiguousarray(images[keep])
    queries = np.ascontiguousarray(coords[qi].astype(qdtype))
    edges = np.ascontiguousarray(np.linspace(0.0, rmax, nbins + 1), dtype=np.float64)
    hist = np.zeros(nbins, dtype=np.int64)
    corc.lib().oracle_rdf_hist(images.ctypes.data, 0, len(images), queries.ctypes.data, int(qdtype == np.float64),
                               0, len(queries), edges.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), nbins,
                               hist.ctypes.data, 0)
    return hist, queries


CASES = [
    # name, generator, dim, side of the query cube / square (>= 32 768 queries)
    ("cfg3_fcc_4M", lambda: cfg3_frame(0), 3, 35.0),
    ("cfg5_uniform_16M", lambda: cfg5_frame(0), 3, 35.0),
    ("cfg2ii_clustered_1M", lambda: cfg2_frame(0, clustered=True), 2, 330.0),
    ("cfg2_uniform_1M", lambda: cfg2_frame(0), 2, 230.0),
]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_large_frame_query_sample_matches_oracle(sc, case):
    """RDF(rmax=10, nbins=500) of one full-size frame: (i) the counts of a >= 32 768-query sample
    equal the brute-force oracle's, float32 and float64 regime; (ii) the self RDF of the full
    frame (symmetric walk) equals the ordered walk and the cross RDF of the frame with itself."""
    from amep_b200 import _lib
    name, gen, dim, side = case
    c64, L = gen()
    rmax, nbins = 10.0, 500
    ctx = _lib.context(0)
    for dtype in (np.float32, np.float64):
        coords = c64.astype(dtype)
        box = (box2d if dim == 2 else box3d)(L, dtype)
        want, queries = _oracle_sample_counts(coords, box, L, side, rmax, nbins, dim, dtype)
        assert len(queries) >= 32768 and want.sum() > 0
        dev = torch.from_numpy(coords).cuda()[None]
        _, _, got = sc.rdf_frames(dev, box, torch.from_numpy(queries).cuda()[None], rmax=rmax, nbins=nbins,
                                  return_counts=True)
        np.testing.assert_array_equal(got[0], want, err_msg=f"{name} {np.dtype(dtype).name}")
        if dtype == np.float32:
            _, _, full = sc.rdf_frames(dev, box, rmax=rmax, nbins=nbins, return_counts=True)
            ctx.set_symmetric(False)
            try:
                _, _, ordered = sc.rdf_frames(dev, box, rmax=rmax, nbins=nbins, return_counts=True)
            finally:
                ctx.set_symmetric(True)
            np.testing.assert_array_equal(full, ordered)
            _, _, cross = sc.rdf_frames(dev, box, dev, rmax=rmax, nbins=nbins, return_counts=True)
            np.testing.assert_array_equal(full, cross)
            n = len(coords)
            if "uniform" in name:   # ideal gas: sum of counts within 6 sigma of N rho V(rmax)
                vol = np.pi * rmax ** 2 if dim == 2 else 4 / 3 * np.pi * rmax ** 3
                expect = n * (n / L ** dim) * vol
                assert abs(int(full.sum()) - expect) < 6 * np.sqrt(2 * expect)
        del dev
        torch.cuda.empty_cache()


def test_intra_frame_shards_add_up(sc):
    """amepb_rdf_set_shard: the histograms of the shares of one frame sum to the full histogram
    exactly (what evaluate.RDF does across GPUs when there are fewer frames than ranks)."""
    from amep_b200 import _core
    for gen, dim in ((lambda: cfg2_frame(1, n=200_000), 2), (lambda: cfg5_frame(1, n=300_000), 3)):
        c64, L = gen()
        coords = torch.from_numpy(c64.astype(np.float32)).cuda()[None]
        box = (box2d if dim == 2 else box3d)(L)
        edges = np.linspace(0.0, 10.0, 501)
        full, _ = _core.rdf_histograms(coords, box, edges)
        for count in (2, 3, 8):
            total = torch.zeros_like(full)
            parts = []
            for index in range(count):
                part, _ = _core.rdf_histograms(coords, box, edges, shard=(index, count))
                parts.append(int(part.sum().item()))
                total += part
            assert torch.equal(total, full)
            assert min(parts) > 0.5 * max(parts)           # interleaved runs: balanced shares
        again, _ = _core.rdf_histograms(coords, box, edges)   # the shard setting does not stick
        assert torch.equal(again, full)


# ------------------------------------------------------------------------------------------
# sf2d in the reference's order: how close to bit-identical is it?
# ------------------------------------------------------------------------------------------
def _ulp_distance_f32(a, b):
    ia = a.astype(np.float32).view(np.int32).astype(np.int64)
    ib = b.astype(np.float32).view(np.int32).astype(np.int64)
    ia = np.where(ia < 0, -(ia & 0x7fffffff), ia)
    ib = np.where(ib < 0, -(ib & 0x7fffffff), ib)
    return np.abs(ia - ib)


def test_sf2d_reference_order_bit_identity(sc, capsys):
    """SURVEY.md A.3 / DESIGN.md section 4: the complex64 running sums of sf2d(mode='std') are
    reproduced in the reference's order; the table values differ from libm's exp by ~1e-16, which
    flips the float32 rounding of a running sum now and then.  Measured here on rho(q) itself
    (before the host epilogue): the fraction of bit-identical components and the largest distance
    in float32 ulps, one chunk and many chunks."""
    c64, L = cfg4_frame(0, n=16384)
    coords = c64.astype(np.float32)
    box = box2d(L)
    qmax = 40.5 * 2 * np.pi / L                # 80 x 80 grid
    for cs in (None, 715, 64):
        cs_used = orc.sf2d_chunksize(len(coords), len(coords), cs)
        want, qx, qy = corc.sf2d_modes(coords, box, qmax, cs_used)
        want = want.reshape(qx.shape)
        s, gx, gy, rho, _ = sc.sf2d_frames(coords[None], box, qmax=qmax, chunksize=cs, return_modes=True)
        np.testing.assert_array_equal(gx[0], qx)
        rho = rho[0]
        comp_got = np.stack([rho.real, rho.imag]).astype(np.float32)
        comp_want = np.stack([want.real, want.imag]).astype(np.float32)
        frac = float(np.mean(comp_got == comp_want))
        ulps = _ulp_distance_f32(comp_got, comp_want)
        # relative to the magnitude of the mode (a component near zero has tiny ulps)
        scale = np.maximum(np.abs(want), 1e-30)
        rel = float(np.max(np.abs(rho - want) / scale))
        with capsys.disabled():
            print(f"\n[sf2d c64 vs reference order] chunksize={cs_used}: identical components {frac:.6f}, "
                  f"max ulp distance {int(ulps.max())}, max |d rho|/|rho| {rel:.3e}")
        assert frac >= 0.999, frac        # measured on B200: 1.000000, 0 ulp, all three chunk sizes
        assert rel < 1e-6
        # S itself (host epilogue = the reference's NumPy expression)
        s_want = np.real(want * want.conj()) / len(coords)
        assert float(np.mean(s[0] == s_want)) >= 0.999
