"""pcf_angle / spatialcor (remaining functions) on the GPU against golden
vectors of the unmodified reference (oracle/make_golden_paircor.py) and the CPU oracle.

Parity bar: integer pair counts equal (distance bins are decided in the reference's own float
arithmetic; angle bins in float64 through arccos, where a pair within an ulp of an angle edge could
in principle differ -- none does in these fixtures); value sums of spatialcor in float64 against the
reference's float32 dot products: 1e-5 relative (float32 inputs), 1e-10 (float64 inputs)."""
import glob
import os

import numpy as np
import pytest

import amep_oracle as orc
from conftest import GOLDEN

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

PCFA = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "pcfa_*.npz")))
SCOR = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "scor_*.npz")))


@pytest.fixture(scope="module")
def sc():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import amep_b200
    return amep_b200.spatialcor


def _dev(a, space):
    return a if (a is None or space == "host") else torch.from_numpy(a).cuda()


def _pcfa_kw(z):
    kw = dict(ndbins=int(z["ndbins"]), nabins=int(z["nabins"]), pbc=bool(z["pbc"]), e=z["e"])
    if not np.isnan(z["rmax"]):
        kw["rmax"] = float(z["rmax"])
    if "psi" in z.files:
        kw["psi"] = z["psi"]
    return kw


@pytest.mark.parametrize("space", ["host", "device"])
@pytest.mark.parametrize("name", PCFA)
def test_pcf_angle_golden(sc, name, space):
    z = np.load(os.path.join(GOLDEN, name))
    coords = _dev(z["coords"].copy(), space)
    other = _dev(z["other"].copy(), space) if "other" in z.files else None
    g, R, T = sc.pcf_angle(coords, z["box"], other_coords=other, **_pcfa_kw(z))
    assert g.dtype == z["g"].dtype and g.shape == z["g"].shape
    np.testing.assert_array_equal(R, z["R"])
    np.testing.assert_array_equal(T, z["T"])
    np.testing.assert_array_equal(g, z["g"])       # counts equal => g bit-identical (same NumPy epilogue)
    assert z["counts"].sum() > 0


def test_pcf_angle_errors_and_side_effects(sc):
    box = np.array([[0, 5], [0, 5], [-0.5, 0.5]])
    c = np.random.default_rng(1).uniform(0, 5, (50, 3))
    with pytest.raises(ValueError):            # not flat
        sc.pcf_angle(c, box, ndbins=5, nabins=4)
    c[:, 2] = 0
    with pytest.raises(ValueError):
        sc.pcf_angle(c, box, ndbins=5, nabins=4, e=np.array([1.0, 0.0]))
    with pytest.raises(ValueError):
        sc.pcf_angle(c, box, ndbins=5, nabins=4, e=np.ones((7, 3)))
    g, R, T = sc.pcf_angle(c, box, ndbins=None, nabins=None, rmax=1.0)
    assert g.shape == (100, 500)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_pcf_angle_larger_against_oracle(sc, dtype):
    rng = np.random.default_rng(81)
    n, m, L = 2500, 900, 30.0
    box = np.array([[-L / 2, L / 2], [0, L], [-0.5, 0.5]], dtype=dtype)
    c = np.zeros((n, 3), dtype=dtype)
    c[:, 0] = rng.uniform(-L / 2 - 1, L / 2 + 1, n)        # a few particles outside the box
    c[:, 1] = rng.uniform(0, L, n)
    o = np.zeros((m, 3), dtype=dtype)
    o[:, 0] = rng.uniform(-L / 2, L / 2, m)
    o[:, 1] = rng.uniform(0, L, m)
    o[:5] = c[:5]                                          # coincident pairs: dist == 0 is dropped
    psi = np.array([-0.2, 0.7])
    for other, kw in ((None, dict(ndbins=60, nabins=30, psi=psi)), (o, dict(ndbins=45, nabins=50, rmax=9.0))):
        want = orc.pcf_angle(c.copy(), box, None if other is None else other.copy(), **kw)
        got = sc.pcf_angle(torch.from_numpy(c.copy()).cuda(), box,
                           other_coords=None if other is None else torch.from_numpy(other.copy()).cuda(), **kw)
        for a, b in zip(got, want):
            np.testing.assert_array_equal(a, b)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_pcf_angle_float32_angle_decisions_equal_float64(sc, monkeypatch, dtype):
    """4e8 pairs: deciding the angle bin in float32 where the 6e-5 rad margin allows it, counting in the
    shared-memory grid, gives the same integers as the float64 arccos of every pair with global atomics
    (AMEPB_PCFA_NO_F32_ANGLE, AMEPB_PCFA_GLOBAL_ATOMICS, AMEPB_PCFA_NO_PRETEST) -- with and without the psi rotation, with a
    reference direction off the x axis, 100 and 2000 angle bins (the margin is 2 % of a bin there)."""
    rng = np.random.default_rng(83)
    n, L = 20000, 100.0
    box = np.array([[0, L], [0, L], [-0.5, 0.5]], dtype=dtype)
    c = np.zeros((n, 3), dtype=dtype)
    c[:, :2] = rng.uniform(0, L, (n, 2))
    ct = torch.from_numpy(c).cuda()
    cases = (dict(), dict(psi=np.array([0.8, 0.3])), dict(e=np.array([0.3, -0.9, 0.0]), nabins=2000, ndbins=20),
             dict(rmax=7.5, ndbins=750))
    for kw in cases:
        out = {}
        for mode in ("fast", "f64"):
            for var in ("AMEPB_PCFA_NO_F32_ANGLE", "AMEPB_PCFA_GLOBAL_ATOMICS", "AMEPB_PCFA_NO_PRETEST"):
                if mode == "f64":
                    monkeypatch.setenv(var, "1")
                else:
                    monkeypatch.delenv(var, raising=False)
            out[mode] = sc.pcf_angle(ct.clone(), box, **kw)
        assert np.isfinite(out["fast"][0]).all() and out["fast"][0].sum() > 0
        for a, b in zip(out["fast"], out["f64"]):
            np.testing.assert_array_equal(a, b)


def _scor_kw(z, name):
    kw = dict(pbc=bool(z["pbc"]))
    if not np.isnan(z["rmax"]):
        kw["rmax"] = float(z["rmax"])
    if int(z["ndists"]) >= 0:
        kw["ndists"] = int(z["ndists"])
    if "nopbc" in name:
        kw["dbin_edges"] = z["edges"]
    return kw


@pytest.mark.parametrize("space", ["host", "device"])
@pytest.mark.parametrize("name", SCOR)
def test_spatialcor_golden(sc, name, space):
    from amep_b200 import _core
    z = np.load(os.path.join(GOLDEN, name))
    coords, values = _dev(z["coords"], space), _dev(z["values"], space)
    other = _dev(z["other"], space) if "other" in z.files else None
    ovals = _dev(z["other_values"], space) if "other" in z.files else None
    kw = _scor_kw(z, name)
    c, r = sc.spatialcor(coords, z["box"], values, other_coords=other, other_values=ovals, **kw)
    np.testing.assert_array_equal(r, z["r"])
    assert c.dtype == z["c"].dtype and c.shape == z["c"].shape
    f32 = z["coords"].dtype == np.float32
    tol = 1e-5 if f32 else 1e-10
    assert np.max(np.abs(c - z["c"])) < tol * max(1.0, float(np.max(np.abs(z["c"]))))
    # the raw sums of the worker: pair counts exact
    norm, cor, _ = _core.spatialcor_sums(coords, values, other, ovals, z["box"], bool(z["pbc"]), z["edges"])
    np.testing.assert_array_equal(norm, z["norm"])
    ref = z["cor"].astype(np.complex128)
    assert np.max(np.abs(cor - ref)) < (2e-5 if f32 else 1e-10) * float(np.max(np.abs(ref)))


def test_spatialcor_larger_against_oracle(sc):
    from amep_b200 import _core
    rng = np.random.default_rng(82)
    n, m = 2600, 1100
    box = np.array([[0, 22], [-5, 12], [0, 9]], dtype=np.float64)
    lo, hi = box[:, 0], box[:, 1]
    c = rng.uniform(lo - 1.0, hi + 1.0, (n, 3))          # unfolded coordinates
    o = rng.uniform(lo, hi, (m, 3))
    v = rng.normal(size=(n, 3))
    w = rng.normal(size=(m, 3))
    for pbc in (True, False):
        edges = orc.spatialcor_edges(box, None, 40, 7.5)
        cor_w, norm_w, _ = orc.spatialcor_sums(c, box, v, o, w, ndists=40, rmax=7.5, pbc=pbc)
        norm, cor, _ = _core.spatialcor_sums(torch.from_numpy(c).cuda(), torch.from_numpy(v).cuda(), torch.from_numpy(o).cuda(),
                                             torch.from_numpy(w).cuda(), box, pbc, edges)
        np.testing.assert_array_equal(norm, norm_w)
        assert np.max(np.abs(cor.real - cor_w)) < 1e-9 * np.max(np.abs(cor_w))
        cg, rg = sc.spatialcor(c, box, v, other_coords=o, other_values=w, ndists=40, rmax=7.5, pbc=pbc)
        cw, rw = orc.spatialcor(c, box, v, o, w, ndists=40, rmax=7.5, pbc=pbc)
        np.testing.assert_array_equal(rg, rw)
        assert np.max(np.abs(cg - cw)) < 1e-9 * np.max(np.abs(cw))


def test_evaluate_pair_classes(sc):
    """PCFangle (three modes), SpatialVelCor, HexOrderCor over a tensor trajectory against the oracle
    composed the way the reference composes them (amep/evaluate.py:1046-1101, 330-338, 2219-2253)."""
    import amep_b200
    z = np.load(os.path.join(GOLDEN, "order_hex_f32.npz"))
    rng = np.random.default_rng(83)
    base = z["coords"]
    nf, n = 3, len(base)
    coords = np.stack([base, base[::-1].copy(), np.roll(base, 7, axis=0)])
    vel = rng.normal(size=(nf, n, 3)).astype(np.float32)
    ang = rng.uniform(0, 2 * np.pi, (nf, n))
    ori = np.stack([np.cos(ang), np.sin(ang), np.zeros_like(ang)], axis=-1)       # float64 directions
    box = z["box"]
    traj = amep_b200.TensorTrajectory(torch.from_numpy(coords).cuda(), box, velocities=torch.from_numpy(vel).cuda(),
                                      orientations=ori)
    kw = dict(ndbins=30, nabins=18, rmax=4.0)
    for mode in ("psi6", "orientations", "x"):
        ev = amep_b200.evaluate.PCFangle(traj, skip=0.0, nav=3, mode=mode, **kw)
        rows = []
        for f in ev.indices:
            psi, e = None, np.array([1.0, 0.0, 0.0])
            if mode == "psi6":
                p = np.mean(orc.psi_k(coords[f], box))
                psi = np.array([p.real, p.imag])
            elif mode == "orientations":
                e = ori[f]
            rows.append(np.array(orc.pcf_angle(coords[f].copy(), box, None, psi=psi, e=e, **kw)))
        want = np.array(rows)
        if mode == "psi6":      # the angle carries the last bits of psi_6: a pair on an angle edge may move
            assert np.mean(ev.frames[:, 0] != want[:, 0]) < 0.02
        else:
            np.testing.assert_array_equal(ev.frames, want)
            np.testing.assert_array_equal(ev.avg, np.mean(want, axis=0)[0])
        np.testing.assert_array_equal(ev.r, np.mean(want, axis=0)[1])
        np.testing.assert_array_equal(ev.theta, np.mean(want, axis=0)[2])
    ev = amep_b200.evaluate.SpatialVelCor(traj, skip=0.0, nav=2, rmax=5.0)
    rows = [np.array(orc.spatialcor(coords[f], box, vel[f], coords[f], vel[f], rmax=5.0)) for f in ev.indices]
    assert np.max(np.abs(ev.frames - np.array(rows))) < 1e-5
    np.testing.assert_array_equal(ev.r, rows[0][1])
    ev = amep_b200.evaluate.HexOrderCor(traj, skip=0.0, nav=2, rmax=5.0)
    rows = [np.array(orc.spatialcor(coords[f], box, orc.psi_k(coords[f], box), rmax=5.0)) for f in ev.indices]
    assert np.max(np.abs(ev.frames - np.array(rows))) < 1e-5
    assert ev.name == "hoc" and ev.avg.shape == rows[0][0].shape


def test_spatialcor_query_slices_add_up(sc):
    """Intra-frame sharding of spatialcor: the pair counts of the slices of the second set add up to
    the unsharded counts exactly, the float64 sums to rounding, and spatialcor(shard=..., reduce=...)
    normalises like the unsharded call (real 3-component values and complex scalars)."""
    from amep_b200 import _core
    rng = np.random.default_rng(85)
    n, L = 6000, 50.0
    box = np.array([[0, L], [0, L], [-0.5, 0.5]], dtype=np.float32)
    c = np.zeros((n, 3), dtype=np.float32)
    c[:, :2] = rng.uniform(0, L, (n, 2))
    ct = torch.from_numpy(c).cuda()
    for values in (torch.from_numpy(rng.normal(size=(n, 3)).astype(np.float32)).cuda(),
                   torch.from_numpy((rng.normal(size=n) + 1j * rng.normal(size=n)).astype(np.complex64)).cuda()):
        want_c, want_r = sc.spatialcor(ct, box, values)
        edges = np.arange(0.0, L / 2, 1.05)
        full_n, full_cor, _ = _core.spatialcor_sums(ct, values, None, None, box, True, edges)
        parts = []
        for index in range(3):
            lo, hi = [(0, 2000), (2000, 2001), (2001, n)][index]
            parts.append(_core.spatialcor_sums(ct, values, ct[lo:hi], values[lo:hi], box, True, edges))
        tot_n = sum(np.asarray(p[0].cpu() if hasattr(p[0], "cpu") else p[0]) for p in parts)
        tot_c = sum(np.asarray(p[1].cpu() if hasattr(p[1], "cpu") else p[1]) for p in parts)
        fn = np.asarray(full_n.cpu() if hasattr(full_n, "cpu") else full_n)
        fc = np.asarray(full_cor.cpu() if hasattr(full_cor, "cpu") else full_cor)
        np.testing.assert_array_equal(tot_n, fn)
        np.testing.assert_allclose(tot_c, fc, rtol=1e-11, atol=1e-9)
        # the public function with an emulated all-reduce over three ranks
        sums = [[], []]

        def collect(a):
            a = np.asarray(a.cpu() if hasattr(a, "cpu") else a)
            sums[len(sums[0]) > len(sums[1])].append(a)
            return a
        for index in range(3):
            sc.spatialcor(ct, box, values, shard=(index, 3), reduce=collect)
        total = [sum(sums[0]), sum(sums[1])]
        it = iter(total)
        got_c, got_r = sc.spatialcor(ct, box, values, shard=(0, 3), reduce=lambda a: next(it))
        np.testing.assert_allclose(got_c, want_c, rtol=1e-9, atol=1e-12)
        np.testing.assert_array_equal(got_r, want_r)


def test_pcf_angle_query_slices_add_up(sc):
    """Intra-frame sharding of pcf_angle (amepb_pcf_angle_set_query_offset): the grids of the query
    slices -- every slice paired with all targets, equal global indices skipped, per-particle
    directions sliced along -- add up to the grid of the whole frame; pcf_angle(shard=..., reduce=...)
    normalises like the unsharded call."""
    from amep_b200 import _core
    rng = np.random.default_rng(87)
    n, L = 5000, 60.0
    box = np.array([[0, L], [0, L], [-0.5, 0.5]], dtype=np.float32)
    c = np.zeros((n, 3), dtype=np.float32)
    c[:, :2] = rng.uniform(0, L, (n, 2))
    c[11] = c[3]                                  # a coincident pair (dist == 0 is dropped either way)
    ct = torch.from_numpy(c).cuda()
    ang = rng.uniform(0, 2 * np.pi, n)
    e_pp = np.stack([np.cos(ang), np.sin(ang), np.zeros(n)], axis=1)
    for e in (np.array([1.0, 0.0, 0.0]), e_pp):
        kw = dict(ndbins=80, nabins=36, rmax=12.0, psi=np.array([0.3, 0.9]), e=e)
        want = sc.pcf_angle(ct.clone(), box, **kw)
        grids = []
        for index in range(3):
            sc.pcf_angle(ct.clone(), box, shard=(index, 3), reduce=lambda a: (grids.append(np.asarray(a)), a)[1], **kw)
        total = sum(grids)
        assert total.sum() > 1e5
        got = sc.pcf_angle(ct.clone(), box, shard=(1, 3), reduce=lambda a: total, **kw)
        for a, b in zip(got, want):
            np.testing.assert_array_equal(a, b)
