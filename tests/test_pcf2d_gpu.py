"""pcf2d parity on the GPU: spatialcor.pcf2d through the C ABI against the reference's golden
vectors and the CPU oracle.  The pair counts are integers: bit-exact; g(x, y), X and Y are the
reference's NumPy expressions on those counts: bit-exact as well."""
import glob
import os
import zlib

import numpy as np
import pytest

import amep_oracle as orc
from conftest import GOLDEN
from test_oracle_golden import pcf2d_kwargs

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

PCF2D_CASES = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "pcf2d_*.npz")))


@pytest.fixture(scope="module")
def sc():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import amep_b200
    return amep_b200.spatialcor


def _rng(tag):
    return np.random.default_rng(zlib.crc32(tag.encode()))


@pytest.fixture(params=["direct", "cells"])
def walk(request, monkeypatch):
    """Both pair walks of the library: the direct kernel (what small systems get) and the
    cell-sorted walk with the shared-memory window, forced here at any size."""
    monkeypatch.setenv("AMEPB_PCF2D_CELLS", "1" if request.param == "cells" else "0")
    return request.param


@pytest.mark.parametrize("space", ["host", "device"])
@pytest.mark.parametrize("name", PCF2D_CASES)
def test_pcf2d_golden(sc, name, space, walk):
    z = np.load(os.path.join(GOLDEN, name))
    coords = z["coords"].copy()
    other = z["other"].copy() if "other" in z.files else None
    if space == "device":
        coords = torch.from_numpy(coords).cuda()
        other = None if other is None else torch.from_numpy(other).cuda()
    g, X, Y = sc.pcf2d(coords, z["box"], other_coords=other, **pcf2d_kwargs(z))
    assert g.dtype == z["g"].dtype and g.shape == z["g"].shape
    np.testing.assert_array_equal(g, z["g"])
    np.testing.assert_array_equal(X, z["X"])
    np.testing.assert_array_equal(Y, z["Y"])
    # the reference flattens the caller's arrays in place (amep/spatialcor.py:846-848)
    last = coords[:, -1].cpu().numpy() if space == "device" else coords[:, -1]
    assert not last.any()


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("rot", [False, True])
def test_pcf2d_counts_vs_oracle_ragged(sc, dtype, rot, walk):
    """Sizes that are no multiple of the query tile (256) or the target stage (1024), several
    target splits, default 500 x 500 grid."""
    from amep_b200 import _core
    rng = _rng(f"pcf2d-ragged-{np.dtype(dtype).name}-{rot}")
    n, m = 2311, 777
    box = np.array([[0, 40], [0, 40], [-0.5, 0.5]], dtype=dtype)
    c = np.zeros((n, 3), dtype=dtype)
    c[:, :2] = rng.uniform(0, 40, (n, 2)).astype(dtype)
    o = np.zeros((m, 3), dtype=dtype)
    o[:, :2] = rng.uniform(0, 40, (m, 2)).astype(dtype)
    psi = np.array([-0.4, 0.6]) if rot else None
    want, xb, yb = orc.pcf2d_counts(c, box, o, psi=psi)
    g, X, Y = sc.pcf2d(c.copy(), box, other_coords=o.copy(), psi=psi)
    gw, Xw, Yw = orc.pcf2d(c, box, o, psi=psi)
    assert want.sum() > 0
    np.testing.assert_array_equal(g, gw)
    np.testing.assert_array_equal(X, Xw)
    # same particles: the equal-index pair is the only one left out
    want_s, _, _ = orc.pcf2d_counts(c, box, None, nxbins=64, nybins=48, rmax=11.0, psi=psi)
    cen = np.mean(box, axis=1)
    r = None
    if rot:
        ang = np.arccos(psi[0] / np.sqrt(psi[0] ** 2 + psi[1] ** 2))
        r = np.array([[1.0, np.cos(-ang), np.sin(-ang)]])
    got_s = _core.pcf2d_counts(torch.from_numpy(c).cuda(), None, cen.astype(np.float64)[None], cen.dtype == np.float32, r,
                               np.linspace(-11.0, 11.0, 65), np.linspace(-11.0, 11.0, 49))
    np.testing.assert_array_equal(got_s[0].cpu().numpy(), want_s)


def test_pcf2d_frames_batch_matches_single_frames(sc, walk):
    """Several frames with their own boxes (own default rmax, own edges) and orientations in one
    library call give what frame-by-frame calls give."""
    rng = _rng("pcf2d-frames")
    nf, n = 3, 900
    boxes = np.array([[[0, 30 + 4 * f], [0, 30 + 4 * f], [-0.5, 0.5]] for f in range(nf)], dtype=np.float32)
    c = np.zeros((nf, n, 3), dtype=np.float32)
    for f in range(nf):
        c[f, :, :2] = rng.uniform(0, 30 + 4 * f, (n, 2)).astype(np.float32)
    psi = np.array([[0.2, 0.9], [1.0, 0.0], [-0.3, -0.8]])
    g, X, Y = sc.pcf2d_frames(torch.from_numpy(c).cuda(), boxes, nxbins=50, nybins=30, psi=psi)
    assert g.shape == (nf, 50, 30) and X.shape == (nf, 30, 50)
    for f in range(nf):
        g1, X1, Y1 = sc.pcf2d(c[f].copy(), boxes[f], nxbins=50, nybins=30, psi=psi[f])
        gw, Xw, Yw = orc.pcf2d(c[f], boxes[f], None, 50, 30, None, psi[f])
        np.testing.assert_array_equal(g[f], g1)
        np.testing.assert_array_equal(g[f], gw)
        np.testing.assert_array_equal(X[f], Xw)
        np.testing.assert_array_equal(Y[f], Yw)


def test_pcf2d_sum_rule_large(sc, walk):
    """Size-independent property at a size the oracle cannot reach: with a window that covers
    every difference vector the counts add up to N*(N-1), and g(x, y) = g(-x, -y) for the same
    particles (every pair is seen from both ends)."""
    from amep_b200 import _core
    rng = _rng("pcf2d-large")
    n = 60000
    c = torch.zeros((n, 3), dtype=torch.float32)
    c[:, :2] = torch.from_numpy(rng.uniform(0, 100, (n, 2)).astype(np.float32))
    edges = np.linspace(-128.0, 128.0, 257)
    counts = _core.pcf2d_counts(c.cuda(), None, np.array([[50.0, 50.0, 0.0]]), True, None, edges, edges)[0].cpu().numpy()
    assert counts.sum() == n * (n - 1)
    # exact float32 differences are antisymmetric, bins are [lo, hi): only differences that sit
    # exactly on an edge land one bin off the mirror image
    mirror = counts[::-1, ::-1]
    assert np.abs(counts - mirror).sum() <= 1e-4 * counts.sum()


def test_pcf2d_invalid_arguments(sc):
    from amep_b200 import _core, _lib
    c = torch.zeros((10, 3), dtype=torch.float32).cuda()
    with pytest.raises(_lib.AmepbError):
        _core.pcf2d_counts(c, None, np.zeros((1, 3)), True, None, np.array([0.0, 1.0, 1.0]), np.array([0.0, 1.0]))
    with pytest.raises(ValueError):
        _core.pcf2d_counts(c, None, np.zeros((1, 3)), True, None, np.array([0.0]), np.array([0.0, 1.0]))


def test_pcf2d_cells_equal_direct_clustered_and_outliers(sc, monkeypatch):
    """The two walks agree bit for bit on a frame that stresses the cell geometry: a dense blob,
    a far outlier (most cells empty), NaN / inf rows, coincident particles, rotation."""
    from amep_b200 import _core
    rng = _rng("pcf2d-cells-clustered")
    n = 6000
    c = np.zeros((n, 3), dtype=np.float32)
    c[:, :2] = rng.uniform(0, 50, (n, 2))
    c[:2000, :2] = rng.normal(25.0, 0.7, (2000, 2))
    c[10, :2] = (400.0, -300.0)        # stretches the cell grid to ~20 x 16 mostly empty cells
    c[11, :2] = (np.nan, 3.0)
    c[12, :2] = (7.0, np.inf)
    c[13] = c[14]
    edges_x, edges_y = np.linspace(-30.0, 30.0, 121), np.linspace(-20.0, 25.0, 91)
    cen = np.array([[25.0, 25.0, 0.0]])
    for rot in (None, np.array([[1.0, np.cos(-1.1), np.sin(-1.1)]])):
        out = {}
        for mode in ("0", "1"):
            monkeypatch.setenv("AMEPB_PCF2D_CELLS", mode)
            out[mode] = _core.pcf2d_counts(torch.from_numpy(c).cuda(), None, cen, True, rot, edges_x, edges_y)[0].cpu().numpy()
        assert out["0"].sum() > 0
        np.testing.assert_array_equal(out["0"], out["1"])


def test_pcf2d_nonuniform_edges(walk):
    """Arbitrary increasing edges through the C ABI (the reference only makes uniform ones): the
    threshold search cannot rely on the spacing; checked against np.histogram2d on the float32
    differences."""
    from amep_b200 import _core
    rng = _rng("pcf2d-nonuniform")
    n = 1500
    c = np.zeros((n, 3), dtype=np.float32)
    c[:, :2] = rng.uniform(0, 20, (n, 2)).astype(np.float32)
    xe = np.concatenate([[-6.0], -6.0 + np.cumsum(rng.uniform(0.01, 0.6, 40))])
    ye = np.concatenate([[-4.0], -4.0 + np.cumsum(rng.uniform(0.05, 0.3, 55))])
    got = _core.pcf2d_counts(torch.from_numpy(c).cuda(), None, np.array([[10.0, 10.0, 0.0]]), True, None, xe, ye)[0].cpu().numpy()
    dx = (c[:, None, 0] - c[None, :, 0])
    dy = (c[:, None, 1] - c[None, :, 1])
    off = ~np.eye(n, dtype=bool)
    want = np.histogram2d(dx[off], dy[off], [xe, ye])[0].astype(np.int64)
    np.testing.assert_array_equal(got, want)


@pytest.mark.parametrize("rot", [False, True])
def test_pcf2d_many_bins_unstaged_edges(sc, rot, walk):
    """More edges than the kernels stage in shared memory (x + y > 2048): bins are searched in
    the global edge arrays."""
    rng = _rng(f"pcf2d-manybins-{rot}")
    n = 900
    box = np.array([[0, 25], [0, 25], [-0.5, 0.5]], dtype=np.float32)
    c = np.zeros((n, 3), dtype=np.float32)
    c[:, :2] = rng.uniform(0, 25, (n, 2)).astype(np.float32)
    psi = np.array([0.6, 0.2]) if rot else None
    g, X, Y = sc.pcf2d(c.copy(), box, nxbins=1100, nybins=1030, rmax=6.0, psi=psi)
    gw, Xw, Yw = orc.pcf2d(c, box, None, 1100, 1030, 6.0, psi)
    np.testing.assert_array_equal(g, gw)
    np.testing.assert_array_equal(Y, Yw)


@pytest.mark.parametrize("box_dtype", [np.float32, np.float64])
def test_pcf2d_rotated_dense_vs_oracle(sc, box_dtype):
    """2e7 binned pairs with rotation on narrow bins: the float32 evaluation that decides most
    rotated pairs must never disagree with the float64 binning of the reference.  (The box is
    centred on the origin: the reference rotates the *difference vectors* about the box centre,
    which moves them out of the window for any other box.)"""
    rng = _rng(f"pcf2d-rot-dense-{np.dtype(box_dtype).name}")
    n = 8000
    box = np.array([[-30, 30], [-30, 30], [-0.5, 0.5]], dtype=box_dtype)
    c = np.zeros((n, 3), dtype=np.float32)
    c[:, :2] = rng.uniform(-30, 30, (n, 2)).astype(np.float32)
    psi = np.array([-0.35, -0.61])
    g, X, Y = sc.pcf2d(c.copy(), box, nxbins=400, nybins=380, rmax=20.0, psi=psi)
    want, _, _ = orc.pcf2d_counts(c, box, None, 400, 380, 20.0, psi)
    gw, _, _ = orc.pcf2d(c, box, None, 400, 380, 20.0, psi)
    assert want.sum() > 1e7
    np.testing.assert_array_equal(g, gw)


@pytest.mark.parametrize("centre", [(0.0, 0.0), (0.37, -0.21)])
def test_pcf2d_rotated_float32_decisions_equal_float64(sc, monkeypatch, centre, walk):
    """1e9 rotated pairs: deciding bins in float32 where the error margin allows gives the same
    counts as sending every pair through the float64 path (AMEPB_PCF2D_NO_F32ROT)."""
    from amep_b200 import _core
    rng = _rng(f"pcf2d-rot-f32-{centre}")
    n = 32000
    c = torch.zeros((n, 3), dtype=torch.float32)
    c[:, :2] = torch.from_numpy(rng.uniform(-40, 40, (n, 2)).astype(np.float32))
    c = c.cuda()
    edges = np.linspace(-25.0, 25.0, 501)
    cen = np.array([[centre[0], centre[1], 0.0]])
    rot = np.array([[1.0, np.cos(-2.2), np.sin(-2.2)]])
    out = {}
    for mode in ("fast", "f64"):
        if mode == "f64":
            monkeypatch.setenv("AMEPB_PCF2D_NO_F32ROT", "1")
        else:
            monkeypatch.delenv("AMEPB_PCF2D_NO_F32ROT", raising=False)
        out[mode] = _core.pcf2d_counts(c, None, cen, True, rot, edges, edges)[0].cpu().numpy()
    assert out["fast"].sum() > 2e8
    np.testing.assert_array_equal(out["fast"], out["f64"])


def test_evaluate_pcf2d(sc):
    """evaluate.PCF2d: frame selection and average of the reference's average_func, orientation
    supplied by the caller (array, callable, or False), loud without it; the trajectory is not
    modified."""
    import amep_b200
    rng = _rng("evaluate-pcf2d")
    nf, n = 6, 400
    c = np.zeros((nf, n, 3), dtype=np.float32)
    c[:, :, :2] = rng.uniform(-8, 8, (nf, n, 2)).astype(np.float32)
    c[:, :, 2] = 0.125
    box = np.array([[-8, 8], [-8, 8], [-0.5, 0.5]], dtype=np.float32)
    psi = rng.normal(size=(nf, 2))
    for dev in ("numpy", "cuda"):
        data = c.copy() if dev == "numpy" else torch.from_numpy(c.copy()).cuda()
        traj = amep_b200.TensorTrajectory(data, box)
        ev = amep_b200.evaluate.PCF2d(traj, skip=0.2, nav=3, psi=psi, nxbins=24, nybins=24, rmax=5.0, njobs=4)
        want = [orc.pcf2d(c[i], box, None, 24, 24, 5.0, psi[i]) for i in ev.indices]
        np.testing.assert_array_equal(ev.avg, np.mean(np.array([w[0] for w in want]), axis=0))
        # x, y are means over the frames too (np.mean(np.array(results), axis=0), amep/utils.py:184-186)
        np.testing.assert_array_equal(ev.x, np.mean(np.array([w[1] for w in want]), axis=0))
        np.testing.assert_array_equal(ev.y, np.mean(np.array([w[2] for w in want]), axis=0))
        np.testing.assert_array_equal(ev.frames[:, 0], np.array([w[0] for w in want]))
        ev2 = amep_b200.evaluate.PCF2d(traj, skip=0.2, nav=3, psi=lambda frame: psi[2], nxbins=24, nybins=24, rmax=5.0)
        want2 = [orc.pcf2d(c[i], box, None, 24, 24, 5.0, psi[2]) for i in ev2.indices]
        np.testing.assert_array_equal(ev2.avg, np.mean(np.array([w[0] for w in want2]), axis=0))
        ev3 = amep_b200.evaluate.PCF2d(traj, skip=0.0, nav=2, psi=False, nxbins=16, nybins=16, rmax=4.0)
        want3 = [orc.pcf2d(c[i], box, None, 16, 16, 4.0, None) for i in ev3.indices]
        np.testing.assert_array_equal(ev3.avg, np.mean(np.array([w[0] for w in want3]), axis=0))
        z = data[:, :, 2].cpu().numpy() if dev == "cuda" else data[:, :, 2]
        assert np.all(z == 0.125)


@pytest.mark.parametrize("walk", ["direct", "cells"])
def test_pcf2d_query_slices_add_up(sc, monkeypatch, walk):
    """Intra-frame sharding (amepb_pcf2d_set_query_offset): the grids of the query slices of a frame --
    every slice paired with all targets, equal global indices skipped -- add up to the grid of the whole
    frame, with and without the rotation, through both pair walks; pcf2d_frames(shard=..., reduce=...)
    then normalises like the unsharded call."""
    from amep_b200 import _core, spatialcor
    monkeypatch.setenv("AMEPB_PCF2D_CELLS", "1" if walk == "cells" else "0")
    rng = np.random.default_rng(91)
    n, L = 30000, 120.0
    c = np.zeros((1, n, 3), dtype=np.float32)
    c[0, :, :2] = rng.uniform(0, L, (n, 2))
    c[0, 17] = c[0, 5]                      # a coincident pair: counted (different indices), unlike a self pair
    ct = torch.from_numpy(c).cuda()
    edges = np.linspace(-30.0, 30.0, 201)[None]
    cen = np.array([[L / 2, L / 2, 0.0]])
    for rot in (None, np.array([[1.0, np.cos(-0.7), np.sin(-0.7)]])):
        full = _core.pcf2d_counts(ct, None, cen, True, rot, edges, edges).cpu().numpy()
        total = np.zeros_like(full)
        for lo, hi in ((0, 9000), (9000, 9001), (9001, 30000)):
            total += _core.pcf2d_counts(ct, ct[:, lo:hi].contiguous(), cen, True, rot, edges, edges, query_offset=lo).cpu().numpy()
        np.testing.assert_array_equal(total, full)
        assert full.sum() > 1e6
    box = np.array([[0, L], [0, L], [-0.5, 0.5]], dtype=np.float32)
    want = spatialcor.pcf2d_frames(ct.clone(), box, nxbins=100, nybins=100, rmax=25.0, psi=np.array([0.6, 0.8]), zero_last_column=False)
    parts = []
    got = None
    for index in range(3):
        acc = []
        g = spatialcor.pcf2d_frames(ct.clone(), box, nxbins=100, nybins=100, rmax=25.0, psi=np.array([0.6, 0.8]),
                                    zero_last_column=False, shard=(index, 3), reduce=lambda a: (acc.append(np.asarray(a.cpu() if hasattr(a, "cpu") else a)), acc[-1])[1])
        parts.append(acc[0])
    total = sum(parts)
    g = spatialcor.pcf2d_frames(ct.clone(), box, nxbins=100, nybins=100, rmax=25.0, psi=np.array([0.6, 0.8]),
                                zero_last_column=False, shard=(0, 3), reduce=lambda a: total)
    for a, b in zip(g, want):
        np.testing.assert_array_equal(a, b)
