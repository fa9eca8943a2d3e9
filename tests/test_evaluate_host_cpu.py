"""Host logic of the evaluate classes on the CPU (no GPU): the kernels are replaced by the
oracle (test infrastructure), what runs is the product's own prologue / epilogue / averaging /
streaming code."""
import threading
import time

import numpy as np
import pytest

import amep_oracle as orc


def _counts_stub(pc, bb, edges, other=None, pbc=True, shard=None, **_):
    """_core.rdf_histograms on the oracle: (int64 [F, nbins], dims [F])."""
    from amep_b200 import _core
    pc = pc if isinstance(pc, _core.Particles) else _core.Particles(pc)
    rows, dims = [], []
    e = np.asarray(edges)
    for f in range(pc.nframes):
        b = bb[f] if np.ndim(bb) == 3 else bb
        ef = e[f] if e.ndim == 2 else e
        o = None if other is None else other.array[f]
        hist, _ = orc.rdf_counts(pc.array[f], b, o, nbins=len(ef) - 1, rmax=float(ef[-1]), pbc=pbc)
        rows.append(hist)
        dims.append(orc.dimension(pc.array[f]))
    return np.array(rows), np.array(dims, dtype=np.int32)


def test_vectorised_rdf_epilogue_equals_the_frame_loop(monkeypatch):
    """rdf_frames normalises all rows of a shared box at once; the reference normalises frame by
    frame (amep/spatialcor.py:639-649).  Same bits, float32 and float64 boxes, 2D and 3D, mixed."""
    from amep_b200 import _core, spatialcor
    monkeypatch.setattr(_core, "rdf_histograms", _counts_stub)
    rng = np.random.default_rng(5)
    for dtype in (np.float32, np.float64):
        box = np.array([[0, 9], [0, 7], [0, 8]], dtype=dtype)
        coords = rng.uniform(0, 7, (5, 60, 3)).astype(dtype)
        coords[1, :, 2] = 0.5          # one 2D frame among 3D ones
        coords[3, :, 2] = 0.5
        for kw in (dict(), dict(rmax=3.0, nbins=40)):
            g, r, counts = spatialcor.rdf_frames(coords, box, return_counts=True, **kw)
            for f in range(5):
                gf, rf = orc.rdf(coords[f], box, **kw)
                assert g[f].dtype == gf.dtype and r[f].dtype == rf.dtype
                np.testing.assert_array_equal(g[f], gf)
                np.testing.assert_array_equal(r[f], rf)


def _sf2d_stub(coords, box, other_coords=None, qmax=20.0, compact_grids=False, **kw):
    coords = np.asarray(coords)
    rows = [orc.sf2d_std(coords[f], box[f] if np.ndim(box) == 3 else box, qmax=qmax) for f in range(coords.shape[0])]
    s = np.array([r[0] for r in rows])
    if compact_grids and np.ndim(box) == 2:
        return s, rows[0][1][None], rows[0][2][None]
    return s, np.array([r[1] for r in rows]), np.array([r[2] for r in rows])


@pytest.mark.parametrize("per_frame_box", [False, True])
def test_sf2d_lazy_frames_and_average_match_average_func(monkeypatch, per_frame_box):
    """SF2d keeps float32 S rows and one pair of grids; ``avg``, ``qx``, ``qy`` and the lazily built
    ``frames`` must equal np.mean(np.array(results), axis=0) / np.array(results) of the reference
    (amep/utils.py:184-186) bit for bit -- including the mean of identical q grids, which is not
    always the grid itself."""
    import amep_b200
    from amep_b200 import evaluate
    monkeypatch.setattr(evaluate, "sf2d_frames", _sf2d_stub)
    rng = np.random.default_rng(9)
    nf, n = 7, 50
    coords = np.zeros((nf, n, 3), dtype=np.float32)
    coords[:, :, :2] = rng.uniform(0, 11, (nf, n, 2))
    box = np.array([[0, 11], [0, 11], [-0.5, 0.5]], dtype=np.float32)
    boxes = box
    if per_frame_box:
        boxes = np.stack([box * (1 + 0.001 * f) for f in range(nf)]).astype(np.float32)
    traj = amep_b200.TensorTrajectory(coords, boxes)
    ev = evaluate.SF2d(traj, skip=0.0, nav=nf, rotate=False, qmax=3.0)
    results = [orc.sf2d_std(coords[i], boxes[i] if per_frame_box else box, qmax=3.0) for i in ev.indices]
    evaluated = np.array(results)
    mean = np.mean(evaluated, axis=0)
    np.testing.assert_array_equal(ev.avg, mean[0])
    np.testing.assert_array_equal(ev.qx, mean[1])
    np.testing.assert_array_equal(ev.qy, mean[2])
    np.testing.assert_array_equal(ev.frames, evaluated)
    assert ev.frames.dtype == np.float64 and ev.avg.dtype == np.float64


class _SlowFrame:
    def __init__(self, traj, i):
        self._t, self._i = traj, i

    def coords(self, ptype=None, **kw):
        self._t.reads.append((threading.current_thread().name, time.perf_counter()))
        time.sleep(self._t.latency)
        c = np.array(self._t.data[self._i])       # a fresh array per call, like an HDF5 read
        return c if ptype is None else c[self._t.types == ptype]

    @property
    def box(self):
        return self._t.box

    @property
    def center(self):
        return np.mean(self._t.box, axis=1)

    def n(self, ptype=None):
        return self._t.data.shape[1]


class _SlowTrajectory:
    """Duck-typed ParticleTrajectory over a memory-mapped file whose frame reads take
    ``latency`` seconds (amep/base.py:941-979: one HDF5 open + read per coords() call)."""

    def __init__(self, path, nf, n, latency):
        rng = np.random.default_rng(3)
        mm = np.lib.format.open_memmap(path, mode="w+", dtype=np.float32, shape=(nf, n, 3))
        mm[:, :, :2] = rng.uniform(0, 12, (nf, n, 2))
        mm[:, :, 2] = 0
        mm.flush()
        self.data = np.load(path, mmap_mode="r")
        self.box = np.array([[0, 12], [0, 12], [-0.5, 0.5]], dtype=np.float32)
        self.types = np.array([1] * (n // 2) + [2] * (n - n // 2))
        self.times = np.arange(nf) * 0.5
        self.latency = latency
        self.reads = []

    def __len__(self):
        return self.data.shape[0]

    def __getitem__(self, item):
        if isinstance(item, (list, tuple, np.ndarray)):
            return [_SlowFrame(self, int(i)) for i in item]
        return _SlowFrame(self, int(item))


def test_streamed_trajectory_overlaps_reads_with_compute(monkeypatch, tmp_path):
    """A trajectory that is not tensor-backed is streamed block by block: a reader thread fills the
    next pinned staging buffer while the current block is computed.  Wall time is about
    max(read, compute), not the sum; the results equal the all-in-memory path bit for bit; at most
    three staging blocks are alive."""
    import amep_b200
    from amep_b200 import _core, evaluate, spatialcor
    nf, n, latency = 24, 64, 0.02
    traj = _SlowTrajectory(str(tmp_path / "traj.npy"), nf, n, latency)
    compute_s = 0.02
    blocks = []

    def slow_counts(pc, bb, edges, other=None, **kw):
        pc = pc if isinstance(pc, _core.Particles) else _core.Particles(pc)
        blocks.append(pc.nframes)
        time.sleep(compute_s * pc.nframes)            # the "GPU" works on this block
        return _counts_stub(pc, bb, edges, other=other, **kw)

    monkeypatch.setattr(_core, "rdf_histograms", slow_counts)
    monkeypatch.setattr(evaluate, "STREAM_BLOCK_BYTES", 4 * n * 12)   # 4 frames per block
    t0 = time.perf_counter()
    ev = evaluate.RDF(traj, skip=0.0, nav=nf, rmax=4.0, nbins=30)
    wall = time.perf_counter() - t0
    read_s, comp_s = nf * latency, nf * compute_s
    assert blocks == [4] * 6
    assert wall < 0.8 * (read_s + comp_s), (wall, read_s, comp_s)         # overlapped ...
    assert wall > max(read_s, comp_s)                                     # ... and honest
    assert {name for name, _ in traj.reads} == {"amep_b200-frame-reader", "MainThread"} or \
        {name for name, _ in traj.reads} == {"amep_b200-frame-reader"}
    # same numbers as the in-memory trajectory
    mem = amep_b200.TensorTrajectory(np.array(traj.data), traj.box, times=traj.times)
    ref = evaluate.RDF(mem, skip=0.0, nav=nf, rmax=4.0, nbins=30)
    np.testing.assert_array_equal(ev.frames, ref.frames)
    np.testing.assert_array_equal(ev.avg, ref.avg)
    np.testing.assert_array_equal(ev.times, ref.times)
    # particle types and a cross RDF through the stream
    evx = evaluate.RDF(traj, skip=0.5, nav=5, ptype=1, other=2, rmax=4.0, nbins=30)
    for k, i in enumerate(evx.indices):
        c = np.array(traj.data[i])
        g, r = orc.rdf(c[traj.types == 1], traj.box, c[traj.types == 2], rmax=4.0, nbins=30)
        np.testing.assert_array_equal(evx.frames[k], np.array([g, r]))


def test_streamed_reader_errors_reach_the_caller(monkeypatch, tmp_path):
    from amep_b200 import _core, evaluate
    traj = _SlowTrajectory(str(tmp_path / "t.npy"), 6, 16, 0.0)
    monkeypatch.setattr(_core, "rdf_histograms", _counts_stub)
    monkeypatch.setattr(evaluate, "STREAM_BLOCK_BYTES", 2 * 16 * 12)
    orig = _SlowFrame.coords

    def broken(self, ptype=None, **kw):
        if self._i == 4:
            raise OSError("disk went away")
        return orig(self, ptype=ptype, **kw)

    monkeypatch.setattr(_SlowFrame, "coords", broken)
    with pytest.raises(OSError, match="disk went away"):
        evaluate.RDF(traj, skip=0.0, nav=6, rmax=4.0, nbins=10)
