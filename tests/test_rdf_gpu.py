"""RDF parity on the GPU: libamepb (through the C ABI) against the reference's
golden vectors and against the CPU oracle on seeded inputs.  Bit-exact."""
import ctypes
import glob
import os
import zlib

import numpy as np
import pytest

import amep_oracle as orc
import c_oracle as corc
from conftest import GOLDEN

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def sc():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import amep_b200
    return amep_b200.spatialcor


RDF_CASES = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "rdf_*.npz"))
                   if "kdtree" not in p)


def _kw(z):
    kw = dict(pbc=bool(z["pbc"]))
    if int(z["nbins"]) >= 0:
        kw["nbins"] = int(z["nbins"])
    if not bool(z["rmax_is_default"]):
        kw["rmax"] = float(z["rmax"])
    return kw


@pytest.mark.parametrize("space", ["host", "device"])
@pytest.mark.parametrize("name", RDF_CASES)
def test_golden_bit_exact(sc, name, space):
    z = np.load(os.path.join(GOLDEN, name))
    coords = z["coords"]
    other = z["other"] if "other" in z.files else None
    if space == "device":
        coords = torch.from_numpy(coords).cuda()
        other = None if other is None else torch.from_numpy(other).cuda()
    g, r, counts = sc.rdf_frames(coords, z["box"], other, return_counts=True, **_kw(z))
    np.testing.assert_array_equal(counts[0], z["counts"])
    g1, r1 = sc.rdf(coords, z["box"], other, **_kw(z))
    assert g1.dtype == z["g"].dtype and r1.dtype == z["r"].dtype
    np.testing.assert_array_equal(g1, z["g"])
    np.testing.assert_array_equal(r1, z["r"])


def test_reference_unittest_criterion(sc):
    """test/test_spatialcor.py:73-102: mode='diff' and mode='kdtree' agree after round(3)."""
    z = np.load(os.path.join(GOLDEN, "rdf_unittest_2d_f64.npz"))
    gk = np.load(os.path.join(GOLDEN, "rdf_unittest_2d_kdtree.npz"))["g"]
    gd, _ = sc.rdf(z["coords"], z["box"], nbins=50, pbc=True, rmax=5.0, mode="diff")
    gt, _ = sc.rdf(z["coords"], z["box"], nbins=50, pbc=True, rmax=5.0, mode="kdtree")
    assert (gd.round(3) == gk.round(3)).all()
    assert (gt.round(3) == gk.round(3)).all()


def _uniform(rng, n, box, dtype, spill=0.0):
    lo, hi = box[:, 0], box[:, 1]
    span = hi - lo
    c = rng.uniform(lo - spill * span, hi + spill * span, (n, 3))
    return c.astype(dtype)


def _clustered(rng, n, box, dtype):
    lo, hi = box[:, 0], box[:, 1]
    span = hi - lo
    centres = rng.uniform(lo, hi, (6, 3))
    half = n // 2
    blob = centres[rng.integers(0, 6, half)] + rng.normal(0, 0.04, (half, 3)) * span
    blob = lo + np.mod(blob - lo, span)
    rest = rng.uniform(lo, hi, (n - half, 3))
    return np.concatenate([blob, rest]).astype(dtype)


CASES = [
    # name, dim, dtype, n, m(other) or None, box, kwargs, generator, spill
    ("2d_f32_smallr", 2, np.float32, 6000, None, [[0, 80], [0, 80], [-0.5, 0.5]], dict(rmax=5.0, nbins=300), "uniform", 0.0),
    ("2d_f64_smallr", 2, np.float64, 6000, None, [[0, 80], [0, 80], [-0.5, 0.5]], dict(rmax=5.0, nbins=300), "uniform", 0.0),
    ("2d_f32_default", 2, np.float32, 3000, None, [[0, 50], [0, 50], [-0.5, 0.5]], dict(), "uniform", 0.0),
    ("2d_f32_clustered", 2, np.float32, 6000, None, [[-40, 40], [-40, 40], [-0.5, 0.5]], dict(rmax=6.0, nbins=500), "clustered", 0.0),
    ("2d_f32_spill", 2, np.float32, 3000, None, [[0, 40], [0, 30], [-0.5, 0.5]], dict(rmax=7.0, nbins=77), "uniform", 0.2),
    ("2d_f64_cross", 2, np.float64, 4000, 1500, [[0, 60], [0, 60], [-0.5, 0.5]], dict(rmax=8.0, nbins=160), "uniform", 0.0),
    ("3d_f32_smallr", 3, np.float32, 6000, None, [[0, 20], [0, 20], [0, 20]], dict(rmax=2.5, nbins=250), "uniform", 0.0),
    ("3d_f64_smallr", 3, np.float64, 5000, None, [[0, 20], [0, 20], [0, 20]], dict(rmax=2.5, nbins=250), "uniform", 0.0),
    ("3d_f32_default", 3, np.float32, 2500, None, [[-5, 5], [-5, 5], [-5, 5]], dict(), "uniform", 0.0),
    ("3d_f32_clustered", 3, np.float32, 6000, None, [[0, 24], [0, 24], [0, 24]], dict(rmax=3.0, nbins=500), "clustered", 0.0),
    ("3d_f32_cross_spill", 3, np.float32, 3000, 2000, [[0, 16], [0, 12], [0, 20]], dict(rmax=4.0, nbins=99), "uniform", 0.1),
    ("3d_f64_thin", 3, np.float64, 3000, None, [[0, 30], [0, 30], [0, 3]], dict(rmax=6.0, nbins=120), "uniform", 0.0),
    ("3d_f32_nopbc", 3, np.float32, 5000, None, [[0, 20], [0, 20], [0, 20]], dict(rmax=3.0, nbins=100, pbc=False), "uniform", 0.0),
    ("3d_f64_nopbc", 3, np.float64, 5000, 1000, [[0, 20], [0, 20], [0, 20]], dict(rmax=3.0, nbins=100, pbc=False), "uniform", 0.0),
]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_oracle_seeded(sc, case):
    name, dim, dtype, n, m, box, kw, gen, spill = case
    rng = np.random.default_rng(zlib.crc32(name.encode()))
    box = np.array(box, dtype=dtype)
    make = _clustered if gen == "clustered" else (lambda r, k, b, d: _uniform(r, k, b, d, spill))
    coords = make(rng, n, box.astype(np.float64), dtype)
    other = None if m is None else make(rng, m, box.astype(np.float64), dtype)
    if dim == 2:
        coords[:, 2] = 0
        if other is not None:
            other[:, 2] = 0
    want, _ = corc.rdf_counts(coords, box, other, **kw)
    g, r, counts = sc.rdf_frames(coords, box, other, return_counts=True, **kw)
    np.testing.assert_array_equal(counts[0], want)
    assert want.sum() > 0
    go, ro = orc.rdf_normalise(want, orc.rdf_edges(box, kw.get("nbins"), kw.get("rmax"))[0], coords, box,
                               n if m is None else m)
    np.testing.assert_array_equal(g[0], go)
    np.testing.assert_array_equal(r[0], ro)


@pytest.mark.parametrize("pbc", [True, False])
@pytest.mark.parametrize("dt_coords,dt_other", [(np.float32, np.float64), (np.float64, np.float32)])
def test_mixed_dtypes_cross(sc, dt_coords, dt_other, pbc):
    """coords and other_coords of different dtypes: the arithmetic regime follows NumPy's promotion of
    ``c - coords`` (amep/spatialcor.py:381), the images stay float32 with pbc."""
    rng = np.random.default_rng(77)
    box = np.array([[0, 25], [0, 25], [0, 25]], dtype=np.float64)
    coords = _uniform(rng, 3000, box, dt_coords)
    other = _uniform(rng, 1200, box, dt_other)
    want, _ = orc.rdf_counts(coords, box, other, rmax=4.0, nbins=120, pbc=pbc)
    _, _, counts = sc.rdf_frames(coords, box, other, rmax=4.0, nbins=120, pbc=pbc, return_counts=True)
    np.testing.assert_array_equal(counts[0], want)
    assert want.sum() > 0


@pytest.mark.parametrize("k", [1, 2, 3, 4, 5, 6])
def test_grid_refinement_does_not_change_counts(sc, k):
    from amep_b200 import _lib
    rng = np.random.default_rng(7)
    box = np.array([[0, 60], [0, 60], [-0.5, 0.5]], dtype=np.float32)
    c = _uniform(rng, 8000, box.astype(np.float64), np.float32)
    c[:, 2] = 0
    want, _ = corc.rdf_counts(c, box, rmax=6.0, nbins=200)
    ctx = _lib.context(0)
    ctx.set_cells_per_cutoff(k)
    try:
        _, _, counts = sc.rdf_frames(c, box, rmax=6.0, nbins=200, return_counts=True)
        info = ctx.rdf_info()
    finally:
        ctx.set_cells_per_cutoff(0)
    np.testing.assert_array_equal(counts[0], want)
    assert 2 * info["candidate_pairs"] >= want.sum()   # symmetric mode evaluates unordered pairs once
    box3 = np.array([[0, 18], [0, 18], [0, 18]], dtype=np.float32)
    c3 = _uniform(rng, 5000, box3.astype(np.float64), np.float32)
    want3, _ = corc.rdf_counts(c3, box3, rmax=3.0, nbins=100)
    ctx.set_cells_per_cutoff(min(k, 4))
    try:
        _, _, counts3 = sc.rdf_frames(c3, box3, rmax=3.0, nbins=100, return_counts=True)
    finally:
        ctx.set_cells_per_cutoff(0)
    np.testing.assert_array_equal(counts3[0], want3)


@pytest.mark.parametrize("sym", [True, False])
def test_dense_cells_query_rounds_and_filter_boundary(sc, sym):
    """Cells with several hundred particles (k = 1): more than one round of 64 staged queries per
    cell, neighbour rows longer than the 256-target staging buffer; plus pairs just below and
    just above the d > 1e-3 filter, with bins narrow enough that the filter sits at 0.4 bin
    widths (the limit of the dump-slot binning is 0.45; beyond it the general kernel runs)."""
    from amep_b200 import _lib
    rng = np.random.default_rng(21)
    box = np.array([[0, 40], [0, 40], [-0.5, 0.5]], dtype=np.float32)
    c = _uniform(rng, 8000, box.astype(np.float64), np.float32)
    c[:, 2] = 0
    for j, d in enumerate((2e-4, 9e-4, 0.99e-3, 1.01e-3, 1.1e-3, 2.4e-3, 2.6e-3)):   # close partners along x
        c[4000 + j, :2] = c[j, :2] + np.float32([d, 0])
    ctx = _lib.context(0)
    ctx.set_symmetric(sym)
    try:
        for k, rmax, nbins in ((1, 8.0, 160), (2, 8.0, 160), (0, 1.25, 500), (0, 1.0, 500)):
            want, _ = corc.rdf_counts(c, box, rmax=rmax, nbins=nbins)
            ctx.set_cells_per_cutoff(k)
            _, _, counts = sc.rdf_frames(c, box, rmax=rmax, nbins=nbins, return_counts=True)
            np.testing.assert_array_equal(counts[0], want, err_msg=f"k={k} rmax={rmax} nbins={nbins}")
    finally:
        ctx.set_cells_per_cutoff(0)
        ctx.set_symmetric(True)
    # 3D, dense: 27 rows of ~190 targets per cell at k = 1
    box3 = np.array([[0, 12], [0, 12], [0, 12]], dtype=np.float32)
    c3 = _uniform(rng, 6000, box3.astype(np.float64), np.float32)
    want3, _ = corc.rdf_counts(c3, box3, rmax=4.0, nbins=100)
    ctx.set_cells_per_cutoff(1)
    try:
        _, _, counts3 = sc.rdf_frames(c3, box3, rmax=4.0, nbins=100, return_counts=True)
    finally:
        ctx.set_cells_per_cutoff(0)
    np.testing.assert_array_equal(counts3[0], want3)


@pytest.mark.parametrize("case", [c for c in CASES if c[4] is None and (c[2] is np.float32 or not c[6].get("pbc", True))],
                         ids=lambda c: c[0])
def test_symmetric_and_ordered_evaluation_agree(sc, case):
    """Self RDFs in the float32 regime are evaluated over unordered pairs (x2); forcing
    every ordered pair must give the same integers, and fewer distance evaluations."""
    from amep_b200 import _lib
    name, dim, dtype, n, m, box, kw, gen, spill = case
    rng = np.random.default_rng(zlib.crc32(name.encode()))
    box = np.array(box, dtype=dtype)
    make = _clustered if gen == "clustered" else (lambda r, k, b, d: _uniform(r, k, b, d, spill))
    coords = make(rng, n, box.astype(np.float64), dtype)
    if dim == 2:
        coords[:, 2] = 0
    ctx = _lib.context(0)
    out = {}
    for mode in (True, False):
        ctx.set_symmetric(mode)
        try:
            _, _, counts = sc.rdf_frames(coords, box, return_counts=True, **kw)
            out[mode] = (counts[0], ctx.rdf_info()["candidate_pairs"])
        finally:
            ctx.set_symmetric(True)
    np.testing.assert_array_equal(out[True][0], out[False][0])
    assert out[True][1] < 0.9 * out[False][1]


@pytest.mark.parametrize("case", [c for c in CASES if c[4] is None and c[2] is np.float32 and c[6].get("pbc", True)],
                         ids=lambda c: c[0])
def test_partitioned_build_equals_placement_kernels(sc, monkeypatch, case):
    """The partitioned build of the query array (strip kernels) and the placement kernels feed the
    pair kernel the same cells: equal integer counts, with 16-byte and with (x, y) records, also for
    particles outside the box, clustered frames and three dimensions (forced for these small
    frames; large batches take it by themselves)."""
    name, dim, dtype, n, m, box, kw, gen, spill = case
    rng = np.random.default_rng(zlib.crc32(name.encode()) + 1)
    box = np.array(box, dtype=dtype)
    make = _clustered if gen == "clustered" else (lambda r, k, b, d: _uniform(r, k, b, d, spill))
    coords = np.stack([make(rng, n, box.astype(np.float64), dtype) for _ in range(3)])
    if dim == 2:
        coords[:, :, 2] = 0
    out = {}
    for mode in ("strips", "strips-vec4", "placement"):
        for var in ("AMEPB_FORCE_PARTITIONED_BUILD", "AMEPB_NO_PARTITIONED_BUILD", "AMEPB_NO_COMPACT_RECORDS"):
            monkeypatch.delenv(var, raising=False)
        if mode == "placement":
            monkeypatch.setenv("AMEPB_NO_PARTITIONED_BUILD", "1")
        else:
            monkeypatch.setenv("AMEPB_FORCE_PARTITIONED_BUILD", "1")
        if mode == "strips-vec4":
            monkeypatch.setenv("AMEPB_NO_COMPACT_RECORDS", "1")
        _, _, counts = sc.rdf_frames(torch.from_numpy(coords).cuda(), box, return_counts=True, **kw)
        out[mode] = counts
    assert out["placement"].sum() > 0
    np.testing.assert_array_equal(out["strips"], out["placement"])
    np.testing.assert_array_equal(out["strips-vec4"], out["placement"])
    want, _ = corc.rdf_counts(coords[1], box, **{k: v for k, v in kw.items() if k in ("nbins", "rmax", "pbc")})
    np.testing.assert_array_equal(out["strips"][1], want)


def test_partitioned_build_with_per_frame_boxes(sc, monkeypatch):
    """Frames with their own boxes (every frame has its own grid, strips and query offset) and a few
    particles outside the box in x: the partitioned build, forced, equals the placement kernels and
    the oracle."""
    rng = np.random.default_rng(12)
    nf, n = 7, 1800
    boxes = np.zeros((nf, 3, 2), dtype=np.float32)
    coords = np.zeros((nf, n, 3), dtype=np.float32)
    for f in range(nf):
        L = 28 + 3 * f
        boxes[f] = [[0, L], [-L / 2, L / 2], [-0.5, 0.5]]
        coords[f, :, 0] = rng.uniform(-1, L + 1, n)
        coords[f, :, 1] = rng.uniform(-L / 2, L / 2, n)
    monkeypatch.setenv("AMEPB_FORCE_PARTITIONED_BUILD", "1")
    g, r, counts = sc.rdf_frames(torch.from_numpy(coords).cuda(), boxes, nbins=150, rmax=9.0, return_counts=True)
    monkeypatch.delenv("AMEPB_FORCE_PARTITIONED_BUILD")
    monkeypatch.setenv("AMEPB_NO_PARTITIONED_BUILD", "1")
    g2, r2, counts2 = sc.rdf_frames(torch.from_numpy(coords).cuda(), boxes, nbins=150, rmax=9.0, return_counts=True)
    np.testing.assert_array_equal(counts, counts2)
    np.testing.assert_array_equal(g, g2)
    for f in (0, 3, 6):
        want, _ = corc.rdf_counts(coords[f], boxes[f], nbins=150, rmax=9.0)
        np.testing.assert_array_equal(counts[f], want)


@pytest.mark.parametrize("seed", range(12))
def test_random_configurations_all_build_paths(sc, monkeypatch, seed):
    """Randomly drawn frames -- dimension, particle number, box shape and offset, cutoff from a few per
    cent of the box to L/2, bin number, batch size, a few particles outside the box -- through every
    build of the cell arrays (partitioned, partitioned with 16-byte records, placement kernels with and
    without (x, y) records): all equal, and equal to the brute-force oracle."""
    rng = np.random.default_rng(9000 + seed)
    dim = 2 if rng.random() < 0.6 else 3
    n = int(rng.integers(300, 5000))
    nf = int(rng.integers(1, 5))
    lens = rng.uniform(8.0, 40.0, 3)
    if dim == 2:
        lens[2] = 1.0
    lo = rng.uniform(-20.0, 20.0, 3)
    if dim == 2:
        lo[2] = -0.5
    box = np.stack([lo, lo + lens], axis=1).astype(np.float32)
    lmin = float(min(lens[:dim]))
    rmax = float(rng.choice([0.07, 0.15, 0.3, 0.5]) * lmin)
    nbins = int(rng.choice([17, 100, 500, 1021]))
    coords = np.zeros((nf, n, 3), dtype=np.float32)
    for d in range(dim):
        coords[:, :, d] = rng.uniform(box[d, 0], box[d, 1], (nf, n))
    if dim == 2:
        coords[:, :, 2] = 0.0
    if rng.random() < 0.3:   # a few particles just outside the box: the ordered walk takes over
        coords[0, :5, 0] = box[0, 1] + rng.uniform(0.0, 0.2, 5)
    out = {}
    for mode, env in (("strips", {"AMEPB_FORCE_PARTITIONED_BUILD": "1"}),
                      ("strips-vec4", {"AMEPB_FORCE_PARTITIONED_BUILD": "1", "AMEPB_NO_COMPACT_RECORDS": "1"}),
                      ("placement", {"AMEPB_NO_PARTITIONED_BUILD": "1"}),
                      ("placement-vec4", {"AMEPB_NO_PARTITIONED_BUILD": "1", "AMEPB_NO_COMPACT_RECORDS": "1"})):
        for var in ("AMEPB_FORCE_PARTITIONED_BUILD", "AMEPB_NO_PARTITIONED_BUILD", "AMEPB_NO_COMPACT_RECORDS"):
            monkeypatch.delenv(var, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        _, _, counts = sc.rdf_frames(torch.from_numpy(coords).cuda(), box, rmax=rmax, nbins=nbins, return_counts=True)
        out[mode] = counts
    for mode in ("strips-vec4", "placement", "placement-vec4"):
        np.testing.assert_array_equal(out[mode], out["strips"], err_msg=f"{mode} seed {seed}")
    f = int(rng.integers(0, nf))
    want, _ = corc.rdf_counts(coords[f], box, nbins=nbins, rmax=rmax)
    np.testing.assert_array_equal(out["strips"][f], want, err_msg=f"oracle seed {seed} dim {dim} n {n} rmax {rmax} nbins {nbins}")


def test_batch_of_frames_with_per_frame_boxes(sc):
    rng = np.random.default_rng(11)
    nf, n = 9, 1500
    boxes = np.zeros((nf, 3, 2), dtype=np.float32)
    coords = np.zeros((nf, n, 3), dtype=np.float32)
    for f in range(nf):
        L = 30 + 2 * f
        boxes[f] = [[0, L], [-L / 2, L / 2], [-0.5, 0.5]]
        coords[f, :, 0] = rng.uniform(0, L, n)
        coords[f, :, 1] = rng.uniform(-L / 2, L / 2, n)
    g, r, counts = sc.rdf_frames(torch.from_numpy(coords).cuda(), boxes, nbins=120, return_counts=True)
    for f in range(nf):
        want, edges = corc.rdf_counts(coords[f], boxes[f], nbins=120)
        np.testing.assert_array_equal(counts[f], want)
        go, ro = orc.rdf_normalise(want, edges, coords[f], boxes[f], n)
        np.testing.assert_array_equal(g[f], go)
        np.testing.assert_array_equal(r[f], ro)


def test_many_small_frames_one_call(sc):
    rng = np.random.default_rng(12)
    nf, n = 64, 400
    box = np.array([[0, 20], [0, 20], [-0.5, 0.5]], dtype=np.float32)
    coords = np.zeros((nf, n, 3), dtype=np.float32)
    coords[:, :, :2] = rng.uniform(0, 20, (nf, n, 2))
    _, _, counts = sc.rdf_frames(coords, box, nbins=64, return_counts=True)
    for f in range(0, nf, 7):
        want, _ = orc.rdf_counts(coords[f], box, nbins=64)
        np.testing.assert_array_equal(counts[f], want)


def test_many_frames_batch_equals_single_frame_calls(sc):
    """Persistent blocks switch frames many times inside one launch; the batch must equal
    frame-by-frame calls (and the host-memory path the device-memory path) bit for bit."""
    g = torch.Generator(device="cuda").manual_seed(21)
    nf, n, L = 96, 20000, 180.0
    c = torch.zeros((nf, n, 3), device="cuda", dtype=torch.float32)
    c[:, :, :2] = torch.rand((nf, n, 2), generator=g, device="cuda") * L
    box = np.array([[0, L], [0, L], [-0.5, 0.5]], dtype=np.float32)
    _, _, batch = sc.rdf_frames(c, box, rmax=10.0, nbins=500, return_counts=True)
    _, _, batch2 = sc.rdf_frames(c, box, rmax=10.0, nbins=500, return_counts=True)
    np.testing.assert_array_equal(batch, batch2)
    host = c.cpu().numpy()
    _, _, via_host = sc.rdf_frames(host, box, rmax=10.0, nbins=500, return_counts=True)
    np.testing.assert_array_equal(batch, via_host)
    for f in (0, 17, 95):
        _, _, one = sc.rdf_frames(c[f:f + 1], box, rmax=10.0, nbins=500, return_counts=True)
        np.testing.assert_array_equal(batch[f], one[0])
    want, _ = corc.rdf_counts(host[40], box, rmax=10.0, nbins=500)
    np.testing.assert_array_equal(batch[40], want)


def test_large_nbins_uses_global_path(sc):
    rng = np.random.default_rng(13)
    box = np.array([[0, 30], [0, 30], [-0.5, 0.5]], dtype=np.float64)
    c = _uniform(rng, 2500, box, np.float64)
    c[:, 2] = 0
    for nbins in (1, 7, 20000):
        want, _ = corc.rdf_counts(c, box, rmax=4.0, nbins=nbins)
        _, _, counts = sc.rdf_frames(c, box, rmax=4.0, nbins=nbins, return_counts=True)
        np.testing.assert_array_equal(counts[0], want)


@pytest.mark.parametrize("space", ["host", "device"])
@pytest.mark.parametrize("cross", [False, True])
def test_sub_batch_pipeline(sc, space, cross, monkeypatch):
    """Many small sub-batches (forced by AMEPB_MAX_FRAMES / AMEPB_STAGE_MB): statistics launched
    one sub-batch ahead, staging buffers reused round robin, per-frame boxes -- every frame must
    equal its single-frame result."""
    from amep_b200 import _core
    rng = np.random.default_rng(55)
    nf, n, m = 23, 1500, 700
    boxes = np.array([[[0, 20 + 0.3 * f], [0, 20 + 0.3 * f], [0, 20 + 0.3 * f]] for f in range(nf)], dtype=np.float32)
    coords = np.stack([_uniform(rng, n, boxes[f].astype(np.float64), np.float32) for f in range(nf)])
    other = np.stack([_uniform(rng, m, boxes[f].astype(np.float64), np.float32) for f in range(nf)]) if cross else None
    coords[7, 3] = np.nan                                   # a frame with a NaN particle in the middle of the batch
    edges = np.linspace(0.0, 3.0, 121)
    single = np.stack([_core.rdf_histograms(coords[f:f + 1], boxes[f], edges,
                                            other=None if other is None else other[f:f + 1])[0][0] for f in range(nf)])
    monkeypatch.setenv("AMEPB_MAX_FRAMES", "3")
    monkeypatch.setenv("AMEPB_STAGE_MB", "1")
    a, b = coords, other
    if space == "device":
        a = torch.from_numpy(coords).cuda()
        b = None if other is None else torch.from_numpy(other).cuda()
    got, dims = _core.rdf_histograms(a, boxes, edges, other=b)
    got = got.cpu().numpy() if torch.is_tensor(got) else got
    np.testing.assert_array_equal(got, single)
    assert (dims == 3).all() and single.sum() > 0


def test_custom_edges_through_the_c_abi(sc):
    """amepb_rdf_batch accepts any non-decreasing edges: a first edge above zero, a first edge below
    the 1e-3 filter, non-uniform edges (general kernel) -- against the C oracle's explicit-edge
    histogram (np.histogram semantics)."""
    from amep_b200 import _core
    rng = np.random.default_rng(91)
    box = np.array([[0, 30], [0, 30], [-0.5, 0.5]], dtype=np.float32)
    c = _uniform(rng, 5000, box.astype(np.float64), np.float32)
    c[:, 2] = 0
    c[2500:2510, :2] = c[:10, :2] + np.float32(4e-4)      # pairs below the filter
    targets = corc.images(c, box)
    def brute(edges):
        e = np.ascontiguousarray(edges, dtype=np.float64)
        hist = np.zeros(len(e) - 1, dtype=np.int64)
        corc.lib().oracle_rdf_hist(targets.ctypes.data, 0, len(targets), c.ctypes.data, 0, 0, len(c),
                                   e.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
                                   len(e) - 1, hist.ctypes.data, 0)
        return hist
    for edges in (np.linspace(1.5, 6.0, 91),                       # uniform, first edge > 0
                  np.linspace(5e-4, 4.0, 401),                     # first edge inside the 1e-3 filter
                  np.linspace(0.0, 0.9, 1001),                     # bins narrower than 1e-3 / 0.45
                  np.concatenate([[0.0, 0.3, 0.35, 1.0], np.geomspace(1.1, 6.0, 40)])):   # non-uniform
        want = brute(edges)
        got, _ = _core.rdf_histograms(c, box, edges)
        np.testing.assert_array_equal(got[0], want)
        got, _ = _core.rdf_histograms(torch.from_numpy(c).cuda(), box, edges)
        np.testing.assert_array_equal(got[0].cpu().numpy(), want)
        assert want.sum() > 0


def test_edge_cases(sc):
    box = np.array([[0, 10], [0, 10], [0, 10]], dtype=np.float64)
    # a single particle is 3D (amep/utils.py:1650-1651); its own images sit at L > rmax
    g, r = sc.rdf(np.array([[1.0, 2.0, 3.0]]), box, nbins=10)
    assert g.shape == (10,) and np.all(g == 0)
    # all particles identical: dimension 0 -> pbc_points raises ValueError (amep/pbc.py:290-294)
    with pytest.raises(ValueError):
        sc.rdf(np.ones((5, 3)), box, nbins=10)
    with pytest.raises(ValueError):
        sc.rdf(np.ones((5, 3)), box, nbins=10, mode="nope")
    with pytest.raises(ValueError):
        sc.rdf(np.ones((5, 2)), box, nbins=10)
    # NaN / inf coordinates never enter a bin, as in np.histogram
    rng = np.random.default_rng(14)
    c = rng.uniform(0, 10, (800, 3))
    c[5] = np.nan
    c[9, 0] = np.inf
    want, _ = orc.rdf_counts(c, box, rmax=3.0, nbins=30)
    _, _, counts = sc.rdf_frames(c, box, rmax=3.0, nbins=30, return_counts=True)
    np.testing.assert_array_equal(counts[0], want)
    # coincident particles: d = 0 is dropped by the d > 1e-3 filter
    c = rng.uniform(0, 10, (500, 3)).astype(np.float32)
    c[100:110] = c[0]
    want, _ = orc.rdf_counts(c, box.astype(np.float32), rmax=2.0, nbins=40)
    _, _, counts = sc.rdf_frames(c, box.astype(np.float32), rmax=2.0, nbins=40, return_counts=True)
    np.testing.assert_array_equal(counts[0], want)
    # bins narrower than the 1e-3 filter
    want, _ = orc.rdf_counts(c, box.astype(np.float32), rmax=0.05, nbins=500)
    _, _, counts = sc.rdf_frames(c, box.astype(np.float32), rmax=0.05, nbins=500, return_counts=True)
    np.testing.assert_array_equal(counts[0], want)


def test_integer_and_noncontiguous_inputs(sc):
    rng = np.random.default_rng(15)
    box = np.array([[0, 40], [0, 40], [0, 40]])
    ci = rng.integers(0, 40, (900, 3))
    want, _ = orc.rdf_counts(ci, box, rmax=5.0, nbins=25)
    _, _, counts = sc.rdf_frames(ci, box, rmax=5.0, nbins=25, return_counts=True)
    np.testing.assert_array_equal(counts[0], want)
    big = rng.uniform(0, 40, (900, 6))
    view = big[:, ::2]
    want, _ = orc.rdf_counts(view, box, rmax=5.0, nbins=25)
    _, _, counts = sc.rdf_frames(view, box, rmax=5.0, nbins=25, return_counts=True)
    np.testing.assert_array_equal(counts[0], want)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64], ids=["f32", "f64"])
def test_full_size_properties(sc, dtype):
    """Size-independent checks at a size the brute-force oracle cannot reach:
    grid refinements agree with each other bit for bit, the total matches the
    ideal-gas expectation, and a random sample of queries matches the oracle."""
    from amep_b200 import _lib
    n, L, rmax = 400_000, 800.0, 10.0
    g = torch.Generator(device="cuda").manual_seed(5)
    c = torch.zeros((1, n, 3), device="cuda", dtype=dtype)
    c[0, :, :2] = (torch.rand((n, 2), generator=g, device="cuda", dtype=torch.float64) * L).to(dtype)
    box = np.array([[0, L], [0, L], [-0.5, 0.5]], dtype=np.float32 if dtype == torch.float32 else np.float64)
    ctx = _lib.context(0)
    results = []
    for k in (1, 2, 3):
        ctx.set_cells_per_cutoff(k)
        try:
            _, _, counts = sc.rdf_frames(c, box, rmax=rmax, nbins=500, return_counts=True)
        finally:
            ctx.set_cells_per_cutoff(0)
        results.append(counts[0])
    np.testing.assert_array_equal(results[0], results[1])
    np.testing.assert_array_equal(results[0], results[2])
    total = results[0].sum()
    expect = n * (n / L ** 2) * np.pi * rmax ** 2
    assert abs(total - expect) < 6 * np.sqrt(expect)
    # cross-RDF of a query sample against the oracle (exact)
    host = c[0].cpu().numpy()
    sample = host[:: n // 200][:200].copy()
    want, _ = corc.rdf_counts(host, box, sample, rmax=rmax, nbins=500)
    _, _, counts = sc.rdf_frames(c, box, torch.from_numpy(sample).cuda()[None], rmax=rmax, nbins=500,
                                 return_counts=True)
    np.testing.assert_array_equal(counts[0], want)


def test_evaluate_rdf_matches_reference(sc):
    import amep_b200
    z = np.load(os.path.join(GOLDEN, "evaluate_rdf_f32.npz"))
    for dev in ("cpu", "cuda"):
        coords = z["coords"] if dev == "cpu" else torch.from_numpy(z["coords"]).cuda()
        traj = amep_b200.TensorTrajectory(coords, z["box"])
        ev = amep_b200.evaluate.RDF(traj, skip=float(z["skip"]), nav=int(z["nav"]), nbins=int(z["nbins"]))
        np.testing.assert_array_equal(ev.indices, z["indices"])
        np.testing.assert_array_equal(ev.frames, z["frames"])
        np.testing.assert_array_equal(ev.avg, z["avg"][0])
        np.testing.assert_array_equal(ev.r, z["avg"][1])
        assert set(["avg", "frames", "indices", "r", "times"]).issubset(ev.keys())
        assert ev["avg"] is ev.avg and ev.name == "rdf"


def test_evaluate_rdf_particle_types_and_cross(sc):
    """evaluate.RDF(ptype=, other=) against the oracle (amep/evaluate.py:513-542)."""
    import amep_b200
    rng = np.random.default_rng(41)
    nf, n, L = 5, 1200, 22.0
    coords = rng.uniform(0, L, (nf, n, 3)).astype(np.float32)
    types = np.array([1] * 700 + [2] * 500)
    box = np.array([[0, L], [0, L], [0, L]], dtype=np.float32)
    for dev in ("cpu", "cuda"):
        data = coords if dev == "cpu" else torch.from_numpy(coords).cuda()
        traj = amep_b200.TensorTrajectory(data, box, types=types)
        ev = amep_b200.evaluate.RDF(traj, skip=0.0, nav=3, ptype=1, other=2, rmax=4.0, nbins=80, max_workers=4, njobs=2)
        idx = orc.frame_indices(nf, 0.0, 3)
        rows = [orc.rdf(coords[i][types == 1], box, coords[i][types == 2], rmax=4.0, nbins=80) for i in idx]
        np.testing.assert_array_equal(ev.frames, np.array(rows))
        np.testing.assert_array_equal(ev.avg, np.mean(np.array(rows), axis=0)[0])


def test_single_target_particle_with_z_images(sc):
    """One target particle is three-dimensional for dimension() (amep/utils.py:1650-1651), so its
    images are shifted along z as well; with z constant in both sets the z-constant fast path must
    not be taken: images at z +- Lz count at sqrt(dxy^2 + Lz^2)."""
    rng = np.random.default_rng(61)
    for dtype in (np.float32, np.float64):
        box = np.array([[0, 9], [0, 9], [0, 4]], dtype=dtype)
        tracer = np.array([[4.5, 4.0, 0.7]], dtype=dtype)          # z off the box centre
        bath = rng.uniform(0, 9, (500, 3)).astype(dtype)
        bath[:, 2] = 0.7                                            # constant z in the query set too
        for rmax in (4.4, 6.5):                                     # below and above Lz
            want, _ = orc.rdf_counts(tracer, box, bath, rmax=rmax, nbins=60)
            _, _, counts = sc.rdf_frames(tracer, box, bath, rmax=rmax, nbins=60, return_counts=True)
            np.testing.assert_array_equal(counts[0], want)
            assert want.sum() > 0


def test_back_to_back_calls_do_not_clobber_the_parameter_block(sc, monkeypatch):
    """The pinned parameter block of a context is shared by asynchronous consumers (plans and
    thresholds of the RDF sub-batches, k / q tables of the S(q) calls).  Device-space calls return
    without synchronising; a following call must not overwrite parameters that are still queued
    behind earlier work.  A long kernel keeps the stream busy while RDF (several
    sub-batches) and sq_harmonics calls with different tables are issued back to back."""
    from amep_b200 import _core
    rng = np.random.default_rng(62)
    nf, n, L = 12, 4000, 30.0
    box = np.array([[0, L], [0, L], [0, L]], dtype=np.float32)
    c = torch.from_numpy(rng.uniform(0, L, (nf, n, 3)).astype(np.float32)).cuda()
    edges_a, edges_b = np.linspace(0.0, 3.0, 101), np.linspace(0.0, 2.0, 81)
    kv_a = np.array([[0.21, 0.0, 0.0], [0.0, 0.4, 0.1]])
    kv_b = np.array([[0.05, 0.3, 0.0], [0.3, 0.1, 0.2]])
    # references: every call alone, synchronised
    ref = {}
    ref["ra"] = _core.rdf_histograms(c, box, edges_a)[0].cpu()
    ref["rb"] = _core.rdf_histograms(c, box, edges_b)[0].cpu()
    ref["ha"] = _core.sq_harmonics(c, kv_a, 40).cpu()
    ref["hb"] = _core.sq_harmonics(c, kv_b, 40).cpu()
    torch.cuda.synchronize()
    monkeypatch.setenv("AMEPB_MAX_FRAMES", "4")          # three sub-batches per RDF call
    busy = torch.empty((8192, 8192), device="cuda")
    for _ in range(3):
        for _ in range(6):
            busy = (busy @ busy).clamp_(-1, 1)           # ~tens of ms of queued work ahead of the calls
        ra = _core.rdf_histograms(c, box, edges_a)[0]
        ha = _core.sq_harmonics(c, kv_a, 40)
        rb = _core.rdf_histograms(c, box, edges_b)[0]
        hb = _core.sq_harmonics(c, kv_b, 40)
        torch.cuda.synchronize()
        assert torch.equal(ra.cpu(), ref["ra"]) and torch.equal(rb.cpu(), ref["rb"])
        assert torch.equal(ha.cpu(), ref["ha"]) and torch.equal(hb.cpu(), ref["hb"])


def test_streamed_trajectory_on_the_gpu(sc, monkeypatch):
    """A trajectory that is not tensor-backed (frame objects with coords() / box, like the
    reference's HDF5 trajectory) is streamed through pinned staging buffers block by block; the
    result equals the in-memory trajectory bit for bit, for RDF and SFiso."""
    import amep_b200
    from amep_b200 import evaluate
    rng = np.random.default_rng(71)
    nf, n, L = 13, 5000, 60.0
    box = np.array([[0, L], [0, L], [-0.5, 0.5]], dtype=np.float32)
    data = np.zeros((nf, n, 3), dtype=np.float32)
    data[:, :, :2] = rng.uniform(0, L, (nf, n, 2))
    types = np.array([1] * 3000 + [2] * 2000)

    class Frame:
        def __init__(self, i):
            self.i = i
            self.box = box

        def coords(self, ptype=None, **kw):
            c = np.array(data[self.i])
            return c if ptype is None else c[types == ptype]

    class Traj:
        times = np.arange(nf) * 2.0

        def __len__(self):
            return nf

        def __getitem__(self, item):
            return [Frame(int(i)) for i in item] if isinstance(item, (list, np.ndarray)) else Frame(int(item))

    monkeypatch.setattr(evaluate, "STREAM_BLOCK_BYTES", 3 * n * 12)      # three frames per block
    mem = amep_b200.TensorTrajectory(data, box, times=Traj.times, types=types)
    for kw in (dict(), dict(ptype=1, other=2)):
        a = evaluate.RDF(Traj(), skip=0.0, nav=nf, rmax=6.0, nbins=120, **kw)
        b = evaluate.RDF(mem, skip=0.0, nav=nf, rmax=6.0, nbins=120, **kw)
        np.testing.assert_array_equal(a.frames, b.frames)
        np.testing.assert_array_equal(a.counts, b.counts)
        np.testing.assert_array_equal(a.times, b.times)
    a = evaluate.SFiso(Traj(), skip=0.2, nav=5, qmax=3.0, num=8)
    b = evaluate.SFiso(mem, skip=0.2, nav=5, qmax=3.0, num=8)
    np.testing.assert_array_equal(a.frames, b.frames)
