"""Pin the CPU oracle (oracle/amep_oracle.py) against outputs of the unmodified
reference stored in tests/golden/ (made by oracle/make_golden.py)."""
import glob
import os

import numpy as np
import pytest

import amep_oracle as orc

from conftest import GOLDEN


def _load(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


RDF_CASES = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "rdf_*.npz"))
                   if "kdtree" not in p)


def rdf_kwargs(z):
    kw = dict(pbc=bool(z["pbc"]))
    if int(z["nbins"]) >= 0:
        kw["nbins"] = int(z["nbins"])
    if not bool(z["rmax_is_default"]):
        kw["rmax"] = float(z["rmax"])
    return kw


@pytest.mark.parametrize("name", RDF_CASES)
def test_rdf_counts_and_g_bit_exact(name):
    z = _load(name)
    other = z["other"] if "other" in z.files else None
    kw = rdf_kwargs(z)
    hist, edges = orc.rdf_counts(z["coords"], z["box"], other, **kw)
    assert hist.dtype == np.int64
    assert edges.dtype == z["edges"].dtype
    np.testing.assert_array_equal(edges, z["edges"])
    np.testing.assert_array_equal(hist, z["counts"])
    g, r = orc.rdf(z["coords"], z["box"], other, **kw)
    assert g.dtype == z["g"].dtype and r.dtype == z["r"].dtype
    np.testing.assert_array_equal(g, z["g"])
    np.testing.assert_array_equal(r, z["r"])


def test_rdf_image_count():
    for name in RDF_CASES:
        z = _load(name)
        if bool(z["pbc"]):
            assert len(orc.periodic_images(z["coords"], z["box"])) == int(z["n_images"])


def test_rdf_reference_unittest_criterion():
    """test/test_spatialcor.py:73-102: diff and kdtree agree to 3 decimals."""
    z = _load("rdf_unittest_2d_f64.npz")
    gk = _load("rdf_unittest_2d_kdtree.npz")["g"]
    g, _ = orc.rdf(z["coords"], z["box"], nbins=50, rmax=5.0)
    assert (g.round(3) == gk.round(3)).all()


@pytest.mark.parametrize("name", ["sfiso_2d_f64.npz", "sfiso_2d_f32.npz", "sfiso_3d_f64.npz",
                                  "sfiso_3d_cross_f32.npz"])
def test_sfiso_fast_bit_exact(name):
    z = _load(name)
    other = z["other"] if "other" in z.files else None
    s, q = orc.sfiso_fast(z["coords"], z["box"], qmax=float(z["qmax"]), twod=bool(z["twod"]),
                          num=int(z["num"]), other_coords=other)
    np.testing.assert_array_equal(q, z["q"])
    np.testing.assert_array_equal(s, z["S"])


FFT_CASES = ["sf2d_fft_f32.npz", "sf2d_fft_f64_cross.npz", "sf2d_fft_f32_offset.npz", "sf2d_fft_lammps_last_f32.npz"]


@pytest.mark.parametrize("name", FFT_CASES)
def test_fft_mode_bit_exact(name):
    """sf2d / sfiso with mode='fft': density histogram, FFT, normalisation, crop, radial average."""
    z = _load(name)
    other = z["other"] if "other" in z.files else None
    ntot = None if int(z["ntot"]) < 0 else int(z["ntot"])
    qmax, acc = float(z["qmax"]), float(z["accuracy"])
    counts, _, _ = orc.density_counts(z["coords"], z["box"], 2 * np.pi / 2 / qmax / (acc * 10))
    np.testing.assert_array_equal(counts, z["counts"])
    S, Qx, Qy = orc.sf2d_fft(z["coords"], z["box"], other, qmax, ntot, acc)
    np.testing.assert_array_equal(S, z["S"])
    np.testing.assert_array_equal(Qx, z["Qx"])
    np.testing.assert_array_equal(Qy, z["Qy"])
    si, qi = orc.sfiso_fft(z["coords"], z["box"], len(z["coords"]), qmax, other, acc)
    np.testing.assert_array_equal(si, z["S_iso"])
    np.testing.assert_array_equal(qi, z["q_iso"])


@pytest.mark.parametrize("name", ["sfiso_std_2d_f64.npz", "sfiso_std_3d_cross_f32.npz", "sfiso_std_2d_pairwise.npz",
                                  "sfiso_std_2d_chunks1000.npz", "sfiso_std_3d_pairwise.npz"])
def test_sfiso_std_debye_bit_exact(name):
    z = _load(name)
    other = z["other"] if "other" in z.files else None
    cs = int(z["chunksize"]) if "chunksize" in z.files else None
    s, q = orc.sfiso_std(z["coords"], z["box"], qmax=float(z["qmax"]), twod=bool(z["twod"]),
                         other_coords=other, chunksize=cs)
    np.testing.assert_array_equal(q, z["q"])
    np.testing.assert_array_equal(s, z["S"])
    # the reference's float32 evaluation sits ~1e-5 from the float64 evaluation of the same sum
    s64, _ = orc.sfiso_std_f64(z["coords"], z["box"], qmax=float(z["qmax"]), twod=bool(z["twod"]),
                               other_coords=other)
    # (2.2e-4 for the 2600-particle frame summed row after row in chunks of 1000)
    assert np.max(np.abs(s64 - s) / np.maximum(np.abs(s64), 1e-2)) < 5e-4


@pytest.mark.parametrize("name", ["sf2d_f64_onechunk.npz", "sf2d_f32_chunks.npz", "sf2d_f64_cross.npz"])
def test_sf2d_std_bit_exact(name):
    z = _load(name)
    other = z["other"] if "other" in z.files else None
    cs = int(z["chunksize"])
    s, qx, qy = orc.sf2d_std(z["coords"], z["box"], other, qmax=float(z["qmax"]),
                             chunksize=None if cs < 0 else cs)
    np.testing.assert_array_equal(qx, z["Qx"])
    np.testing.assert_array_equal(qy, z["Qy"])
    assert s.dtype == z["S"].dtype
    np.testing.assert_array_equal(s, z["S"])
    # the float64 evaluation of the same formula sits within the reference's
    # own complex64 noise
    s64, _, _ = orc.sf2d_f64(z["coords"], z["box"], other, qmax=float(z["qmax"]))
    assert np.max(np.abs(s64 - s)) < 2e-3 * max(1.0, np.max(np.abs(s64)))


def test_evaluate_rdf_matches_reference_average_func():
    z = _load("evaluate_rdf_f32.npz")
    stack, box = z["coords"], z["box"]
    frames = [(stack[i], box) for i in range(len(stack))]
    res, avg, idx = orc.evaluate(frames, lambda f: orc.rdf(f[0], f[1], nbins=int(z["nbins"])),
                                 skip=float(z["skip"]), nav=int(z["nav"]))
    np.testing.assert_array_equal(idx, z["indices"])
    np.testing.assert_array_equal(res, z["frames"])
    np.testing.assert_array_equal(avg, z["avg"])


def test_frame_indices_rule():
    np.testing.assert_array_equal(orc.frame_indices(100, 0.0, 100), np.arange(100))
    np.testing.assert_array_equal(orc.frame_indices(51, 0.9, 2), np.array([46, 50]))
    assert len(orc.frame_indices(10, 0.5, 50)) == 5
    assert len(orc.frame_indices(10, 0.0, None)) == 10


# ---- the C restatement (oracle/rdf_oracle.c) against the same golden vectors -------------
import c_oracle as corc  # noqa: E402


@pytest.mark.parametrize("name", RDF_CASES)
def test_c_oracle_rdf_counts_bit_exact(name):
    z = _load(name)
    other = z["other"] if "other" in z.files else None
    hist, edges = corc.rdf_counts(z["coords"], z["box"], other, **rdf_kwargs(z))
    np.testing.assert_array_equal(hist, z["counts"])
    hist1, _ = corc.rdf_counts(z["coords"], z["box"], other, nthreads=1, **rdf_kwargs(z))
    np.testing.assert_array_equal(hist1, z["counts"])


@pytest.mark.parametrize("name", ["sfiso_2d_f64.npz", "sfiso_2d_f32.npz", "sfiso_3d_f64.npz",
                                  "sfiso_3d_cross_f32.npz"])
def test_c_oracle_sfiso(name):
    z = _load(name)
    other = z["other"] if "other" in z.files else None
    s, q = corc.sfiso_fast(z["coords"], z["box"], qmax=float(z["qmax"]), twod=bool(z["twod"]),
                           num=int(z["num"]), other_coords=other)
    np.testing.assert_array_equal(q, z["q"])
    # plain vs pairwise summation order: 1e-12 relative, not bit-exact
    np.testing.assert_allclose(s, z["S"], rtol=1e-10, atol=1e-12)


@pytest.mark.parametrize("name", ["sf2d_f64_onechunk.npz", "sf2d_f32_chunks.npz", "sf2d_f64_cross.npz"])
def test_c_oracle_sf2d_bit_exact(name):
    z = _load(name)
    other = z["other"] if "other" in z.files else None
    cs = int(z["chunksize"])
    s, qx, qy = corc.sf2d_std(z["coords"], z["box"], other, qmax=float(z["qmax"]),
                              chunksize=None if cs < 0 else cs)
    np.testing.assert_array_equal(s, z["S"])


PCF2D_CASES = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "pcf2d_*.npz")))


def pcf2d_kwargs(z):
    kw = {}
    if int(z["nxbins"]) >= 0:
        kw["nxbins"] = int(z["nxbins"])
    if int(z["nybins"]) >= 0:
        kw["nybins"] = int(z["nybins"])
    if not np.isnan(float(z["rmax"])):
        kw["rmax"] = float(z["rmax"])
    if "psi" in z.files:
        kw["psi"] = z["psi"]
    return kw


@pytest.mark.parametrize("name", PCF2D_CASES)
def test_pcf2d_bit_exact(name):
    """pcf2d: integer pair counts, g(x, y) and the bin-centre grids, bit for bit."""
    assert len(PCF2D_CASES) == 6
    z = _load(name)
    other = z["other"] if "other" in z.files else None
    kw = pcf2d_kwargs(z)
    counts, _, _ = orc.pcf2d_counts(z["coords"], z["box"], other, **kw)
    np.testing.assert_array_equal(counts, z["counts"])
    g, X, Y = orc.pcf2d(z["coords"], z["box"], other, **kw)
    assert g.dtype == z["g"].dtype
    np.testing.assert_array_equal(g, z["g"])
    np.testing.assert_array_equal(X, z["X"])
    np.testing.assert_array_equal(Y, z["Y"])


# ---- round 2: psi_6 prologue, pcf_angle, spatialcor (oracle pinned on reference goldens) -------------
ORDER_CASES = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "order_*.npz")))
PCFA_CASES = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "pcfa_*.npz")))
SCOR_CASES = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "scor_*.npz")))


@pytest.mark.parametrize("name", ORDER_CASES)
def test_oracle_psi6_prologue(name):
    """knn ids, psi_k, mean, angle bit for bit; the rotated crop to 1e-13 (the reference's np.dot may
    fuse a multiply-add in the 3x3 rotation)."""
    z = np.load(os.path.join(GOLDEN, name))
    np.testing.assert_array_equal(orc.knn_ids(z["coords"], z["box"]), z["neighbors"])
    psi = orc.psi_k(z["coords"], z["box"])
    assert psi.dtype == z["psi"].dtype
    np.testing.assert_array_equal(psi, z["psi"])
    assert np.mean(psi) == z["psi_mean"][0]
    assert orc.orientation_angle(np.mean(psi)) == float(z["theta"])
    rc = orc.rotated_periodic_crop(z["coords"], z["box"], float(z["theta"]))
    assert rc.shape == z["rotated"].shape and np.max(np.abs(rc - z["rotated"])) < 1e-13


@pytest.mark.parametrize("name", PCFA_CASES)
def test_oracle_pcf_angle(name):
    z = np.load(os.path.join(GOLDEN, name))
    kw = dict(ndbins=int(z["ndbins"]), nabins=int(z["nabins"]), rmax=None if np.isnan(z["rmax"]) else float(z["rmax"]),
              psi=z["psi"] if "psi" in z.files else None, pbc=bool(z["pbc"]), e=z["e"])
    other = z["other"] if "other" in z.files else None
    counts, _, _ = orc.pcf_angle_counts(z["coords"], z["box"], other, **kw)
    np.testing.assert_array_equal(counts, z["counts"])
    g, R, T = orc.pcf_angle(z["coords"], z["box"], other, **kw)
    np.testing.assert_array_equal(g, z["g"])
    np.testing.assert_array_equal(R, z["R"])
    np.testing.assert_array_equal(T, z["T"])


@pytest.mark.parametrize("name", SCOR_CASES)
def test_oracle_spatialcor(name):
    """Pair counts per bin exact (the periodic KD-tree's distances restated); the value sums are
    float64 here, float32 dot products in the reference: 1e-5 relative."""
    z = np.load(os.path.join(GOLDEN, name))
    kw = dict(rmax=None if np.isnan(z["rmax"]) else float(z["rmax"]), ndists=None if int(z["ndists"]) < 0 else int(z["ndists"]),
              pbc=bool(z["pbc"]))
    if "nopbc" in name:
        kw["dbin_edges"] = z["edges"]
    other = z["other"] if "other" in z.files else None
    ovals = z["other_values"] if "other" in z.files else None
    cor, norm, edges = orc.spatialcor_sums(z["coords"], z["box"], z["values"], other, ovals, **kw)
    np.testing.assert_array_equal(norm, z["norm"])
    np.testing.assert_array_equal(edges, z["edges"])
    assert np.max(np.abs(cor - z["cor"])) < 1e-5 * np.max(np.abs(z["cor"]))
    c, r = orc.spatialcor(z["coords"], z["box"], z["values"], other, ovals, **kw)
    np.testing.assert_array_equal(r, z["r"])
    assert np.max(np.abs(c - z["c"])) < 1e-5
