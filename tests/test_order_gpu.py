"""The psi_6 orientation prologue on the GPU (amepb_psi_k, amepb_rotate_crop) against golden
vectors of the unmodified reference (oracle/make_golden_order.py) and the CPU oracle.

Parity bar: (i) neighbour ids equal to the reference's; (ii) mean psi_6 within
1e-6 absolute; (iii) with the reference's angle injected, SF2d(rotate=True) within 1e-6 like
rotate=False.  The end-to-end default call is also compared, at the accuracy the reference's own
float32 psi_6 allows."""
import glob
import os

import numpy as np
import pytest

import amep_oracle as orc
from conftest import GOLDEN

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

CASES = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "order_*.npz")))


@pytest.fixture(scope="module")
def ab():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import amep_b200
    return amep_b200


def _close(got, want, rtol=1e-6, atol_scale=1e-7):
    atol = atol_scale * float(np.mean(np.abs(want)))
    err = np.abs(got - want)
    assert np.all(err <= rtol * np.abs(want) + atol), f"max rel err {np.max(err / np.maximum(np.abs(want), atol)):.3e}"


@pytest.mark.parametrize("space", ["host", "device"])
@pytest.mark.parametrize("name", CASES)
def test_psi6_golden(ab, name, space):
    from amep_b200 import _core, order
    z = np.load(os.path.join(GOLDEN, name))
    coords = z["coords"] if space == "host" else torch.from_numpy(z["coords"]).cuda()
    psi, ids = _core.psi_k(coords, z["box"], k=6, want_neighbors=True)
    np.testing.assert_array_equal(ids, z["neighbors"])                      # (i)
    assert psi.dtype == z["psi"].dtype
    tol = 2e-6 if psi.dtype == np.complex64 else 1e-12
    assert np.max(np.abs(psi - z["psi"])) < tol
    mean = order.psi6_mean(coords, z["box"])
    want = z["psi_mean"][0]
    assert abs(mean[0] - want.real) < 1e-6 and abs(mean[1] - want.imag) < 1e-6      # (ii)
    theta = order.psi6_orientation(coords, z["box"])
    assert abs(theta - float(z["theta"])) < 1e-4 / max(abs(want), 1e-3)
    # module-level function with the reference's signature
    np.testing.assert_array_equal(order.psi_k(coords, z["box"], k=6), psi)


@pytest.mark.parametrize("space", ["host", "device"])
@pytest.mark.parametrize("name", CASES)
def test_rotated_crop_golden(ab, name, space):
    from amep_b200 import order
    z = np.load(os.path.join(GOLDEN, name))
    theta = float(z["theta"])
    for key_in, key_out in (("coords", "rotated"), ("other", "rotated_other")):
        if key_in not in z.files:
            continue
        c = z[key_in] if space == "host" else torch.from_numpy(z[key_in]).cuda()
        got = order.rotated_periodic_crop(c, z["box"], theta)
        got = got.cpu().numpy() if torch.is_tensor(got) else got
        assert got.dtype == np.float64 and got.shape == z[key_out].shape
        assert np.max(np.abs(got - z[key_out])) < 1e-12      # the reference's np.dot may fuse a multiply-add
        np.testing.assert_array_equal(got[:, 2], z[key_out][:, 2])


@pytest.mark.parametrize("name", CASES)
def test_sf2d_rotate_with_reference_angle(ab, name):
    """(iii): SF2d(rotate=True) with the reference's angle injected -> 1e-6 like rotate=False."""
    z = np.load(os.path.join(GOLDEN, name))
    n = len(z["coords"])
    types = None
    coords = z["coords"]
    kw = {}
    if "other" in z.files:
        coords = np.concatenate([z["coords"], z["other"]])
        types = np.array([1] * n + [2] * len(z["other"]))
        kw = dict(ptype=1, other=2)
    for data in (coords, torch.from_numpy(coords).cuda()):
        traj = ab.TensorTrajectory(data[None], z["box"], types=types)
        ev = ab.evaluate.SF2d(traj, skip=0.0, nav=1, qmax=float(z["qmax"]), orientation=[float(z["theta"])], **kw)
        np.testing.assert_array_equal(ev.qx, z["Qx"])
        np.testing.assert_array_equal(ev.qy, z["Qy"])
        _close(ev.avg, z["S"].astype(np.float64))
        assert ev.frames.shape == (1, 3) + z["S"].shape


@pytest.mark.parametrize("name", [c for c in CASES if "other" not in np.load(os.path.join(GOLDEN, c)).files])
def test_sf2d_default_call_end_to_end(ab, name, capsys):
    """The reference's default call, SF2d(traj): psi_6, angle, rotation, crop and S(q) all on this
    side.  The angle inherits the float32 rounding of psi_6 (ours and the reference's differ in
    the last bits of arctan2/exp), and S(q) moves by q R dtheta: compared at 1e-3 relative to the
    mean level, with the measured deviation printed."""
    z = np.load(os.path.join(GOLDEN, name))
    traj = ab.TensorTrajectory(torch.from_numpy(z["coords"]).cuda()[None], z["box"])
    ev = ab.evaluate.SF2d(traj, skip=0.0, nav=1, qmax=float(z["qmax"]))
    want = z["S"].astype(np.float64)
    dev = float(np.max(np.abs(ev.avg - want)) / np.mean(np.abs(want)))
    with capsys.disabled():
        print(f"\n[SF2d default, {name}] max |dS| / mean|S| = {dev:.3e}")
    assert dev < 1e-3


def test_psi6_larger_against_oracle(ab):
    from amep_b200 import _core
    rng = np.random.default_rng(77)
    n, L = 3000, 50.0
    for dtype in (np.float32, np.float64):
        c = np.zeros((n, 3), dtype=dtype)
        c[:, :2] = rng.uniform(-3.0, L + 2.0, (n, 2)).astype(dtype)          # some particles outside the box
        c[: n // 3, :2] = (rng.normal(0, 2.0, (n // 3, 2)) + 30.0).astype(dtype)   # a dense blob: uneven cell occupancy
        c[:, 2] = rng.uniform(-0.2, 0.2, n).astype(dtype) if dtype == np.float64 else 0    # z enters the distances
        box = np.array([[0, L], [0, L], [-0.5, 0.5]], dtype=dtype)
        want_ids = orc.knn_ids(c, box)
        want_psi = orc.psi_k(c, box)
        psi, ids = _core.psi_k(torch.from_numpy(c).cuda(), box, k=6, want_neighbors=True)
        np.testing.assert_array_equal(ids, want_ids)
        assert np.max(np.abs(psi - want_psi)) < (2e-6 if dtype == np.float32 else 1e-12)
        assert abs(np.mean(psi) - np.mean(want_psi)) < 1e-6


def test_psi6_raises_like_the_reference_for_tiny_systems(ab):
    from amep_b200 import order
    box = np.array([[0, 4], [0, 4], [-0.5, 0.5]])
    c = np.array([[1.0, 1.0, 0], [2.5, 1.2, 0]])   # two particles: their own images rank among the 7 nearest
    with pytest.raises(ValueError):
        orc.knn_ids(c, box)
    with pytest.raises(ValueError):
        order.psi_k(c, box, k=6)
    with pytest.raises(NotImplementedError):
        order.psi_k(c, box, other_coords=c)


def test_pcf2d_default_orientation(ab):
    """evaluate.PCF2d without psi: psi_6 of all particles of the frame, like amep/evaluate.py:730-735."""
    z = np.load(os.path.join(GOLDEN, "order_hex_f32.npz"))
    coords = np.stack([z["coords"], z["coords"][::-1].copy()])
    traj = ab.TensorTrajectory(torch.from_numpy(coords).cuda(), z["box"])
    ev = ab.evaluate.PCF2d(traj, skip=0.0, nav=2, nxbins=40, nybins=40, rmax=4.0)
    psi = [np.mean(orc.psi_k(coords[f], z["box"])) for f in range(2)]
    want = [orc.pcf2d(coords[f].copy(), z["box"], None, 40, 40, 4.0, np.array([p.real, p.imag])) for f, p in enumerate(psi)]
    got = ev.frames
    # counts near a bin edge may move with the last bits of the angle: compare the bulk
    for f in range(2):
        diff = np.abs(got[f, 0] - want[f][0])
        assert np.mean(diff > 0) < 0.05
        np.testing.assert_array_equal(got[f, 1], want[f][1])
