"""world_size-2 `gloo` tests of the frame sharding + all-reduce host logic (no GPU).

The per-frame kernels are replaced by the CPU oracle (test infrastructure) so that
what is exercised is the product's own sharding, table assembly and averaging code."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT

torch = pytest.importorskip("torch")
import torch.multiprocessing as mp  # noqa: E402


def orc_pcf2d(coords, box, psi):
    import amep_oracle as orc
    return orc.pcf2d(coords, box, None, 12, 12, 3.0, psi)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _oracle_rdf_frames(coords, box, other_coords=None, nbins=None, rmax=None, mode="diff", pbc=True,
                       return_counts=False, reduce=None, shard=None):
    """Stand-in for spatialcor.rdf_frames on the CPU oracle.  ``shard=(i, c)`` keeps every c-th query
    (the library deals out query cell runs instead; either way the shares add up to the full
    histogram), ``reduce`` is applied to the counts before the normalisation."""
    import amep_oracle as orc
    coords = np.asarray(coords)
    h_rows, e_rows = [], []
    for f in range(coords.shape[0]):
        b = box[f] if np.ndim(box) == 3 else box
        o = None if other_coords is None else np.asarray(other_coords)[f]
        if shard is not None:
            o = (coords[f] if o is None else o)[shard[0]::shard[1]]
        hist, edges = orc.rdf_counts(coords[f], b, o, nbins=nbins, rmax=rmax, pbc=pbc)
        h_rows.append(hist), e_rows.append(edges)
    hists = np.array(h_rows)
    if reduce is not None:
        hists = reduce(hists)
    g_rows, r_rows = [], []
    for f in range(coords.shape[0]):
        b = box[f] if np.ndim(box) == 3 else box
        nq = len(coords[f]) if other_coords is None else len(np.asarray(other_coords)[f])
        g, r = orc.rdf_normalise(hists[f], e_rows[f], coords[f], b, nq)
        g_rows.append(g), r_rows.append(r)
    out = (np.array(g_rows), np.array(r_rows))
    return out + (hists,) if return_counts else out


def _oracle_sfiso_frames(coords, box, N=None, qmax=20.0, twod=True, mode="fast", other_coords=None,
                         accuracy=0.5, num=8, reduce=None, counts=None, chunksize=None, njobs=1):
    """Stand-in for spatialcor.sfiso_frames(mode='fast') in NumPy: per direction the (cos, sin) sums
    of the particles it is given, ``reduce`` over the ranks, the reference's normalisation
    (amep/spatialcor.py:1440-1446, 1660-1664) with the particle numbers of the whole frame."""
    import amep_oracle as orc
    coords = np.asarray(coords)
    q = orc.sf_qvalues(box, qmax)
    dirs = orc.sfiso_directions(twod, num)
    ha = np.zeros((coords.shape[0], len(dirs), len(q), 2))
    for f in range(coords.shape[0]):
        for d, u in enumerate(dirs):
            proj = coords[f].astype(np.float64) @ u
            ph = q[:, None] * proj[None, :]
            ha[f, d, :, 0], ha[f, d, :, 1] = np.cos(ph).sum(axis=1), np.sin(ph).sum(axis=1)
    if reduce is not None:
        ha = reduce(ha)
    na = coords.shape[1] if counts is None else counts[0]
    s = ((ha[..., 0] ** 2 + ha[..., 1] ** 2) / np.sqrt(na * na)).mean(axis=1)
    return s, np.broadcast_to(q, (coords.shape[0], len(q))).copy()


def _oracle_sf2d_frames(coords, box, other_coords=None, qmax=20.0, compact_grids=False, shard=None, reduce=None, **kw):
    """Stand-in for spatialcor.sf2d_frames(mode='std').  With ``shard`` / ``reduce`` (fewer frames than
    ranks) it does what amepb_sf2d_set_shard does: this rank's rows of q tiles of rho, zeros elsewhere,
    summed over the ranks."""
    import amep_oracle as orc
    coords = np.asarray(coords)
    rows = [orc.sf2d_std(coords[f], box, qmax=qmax) for f in range(coords.shape[0])]
    s = np.array([r[0] for r in rows])
    if shard is not None:
        index, count = shard
        qx, qy = orc.sf2d_grid(box, qmax)
        cs = orc.sf2d_chunksize(coords.shape[1], coords.shape[1])
        rho = np.array([getattr(orc, "_density_modes_c64")(coords[f], qx, qy, cs) for f in range(coords.shape[0])])
        nq = rho.shape[1] // 2
        mine = np.zeros(2 * nq, dtype=bool)
        for b in range(nq):                       # row b of the positive quadrant and its mirror row
            if (b // 16) % count == index:
                mine[nq + b] = mine[nq - 1 - b] = True
        rho = np.where(mine[None, :, None], rho, 0).astype(np.complex64)
        rho = reduce(rho)
        s = np.real(rho * rho.conj()) / coords.shape[1]
    if compact_grids:
        return s, rows[0][1][None], rows[0][2][None]
    return s, np.array([r[1] for r in rows]), np.array([r[2] for r in rows])


def _oracle_pcf2d_frames(coords, box, other_coords=None, nxbins=None, nybins=None, rmax=None, psi=None,
                         zero_last_column=True, shard=None, reduce=None):
    """(``shard`` / ``reduce`` -- fewer frames than ranks -- are accepted and not emulated: every rank
    computes the whole frame here; the slicing itself is tested on the GPU, tests/test_pcf2d_gpu.py)"""
    import amep_oracle as orc
    coords = np.asarray(coords)
    out = []
    for f in range(coords.shape[0]):
        b = box[f] if np.ndim(box) == 3 else box
        o = None if other_coords is None else np.asarray(other_coords)[f]
        out.append(orc.pcf2d(coords[f], b, o, nxbins, nybins, rmax, None if psi is None else psi[f]))
    return tuple(np.array([r[k] for r in out]) for k in range(3))


def _worker(rank, world, port, tmpdir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import amep_b200
    from amep_b200 import dist as adist, evaluate

    # (1) assemble_rows: integer rows and complex rows
    lo, hi = adist.shard_bounds(7, rank, world)
    rows = np.arange(7 * 5, dtype=np.int64).reshape(7, 5) * 3 + 1
    table = adist.assemble_rows(rows[lo:hi], lo, 7)
    assert np.array_equal(table, rows)
    crow = (np.arange(7 * 4).reshape(7, 4) * (1 + 2j)).astype(np.complex64)
    ctab = adist.assemble_rows(crow[lo:hi], lo, 7)
    assert np.array_equal(ctab, crow)

    # (2) evaluate.RDF over a sharded trajectory equals the reference's average_func result
    evaluate.rdf_frames = _oracle_rdf_frames
    z = np.load(os.path.join(GOLDEN, "evaluate_rdf_f32.npz"))
    traj = amep_b200.TensorTrajectory(z["coords"], z["box"])
    ev = evaluate.RDF(traj, skip=float(z["skip"]), nav=int(z["nav"]), nbins=int(z["nbins"]))
    assert np.array_equal(ev.indices, z["indices"])
    assert np.array_equal(ev.frames, z["frames"])
    assert np.array_equal(ev.avg, z["avg"][0]) and np.array_equal(ev.r, z["avg"][1])
    # fewer frames than ranks (nav=1): the frame is sharded *inside* (every rank a share of the
    # queries, integer histograms summed before the normalisation) -- same bits as one rank
    ev1 = evaluate.RDF(traj, skip=0.0, nav=1, nbins=50)
    solo, _, solo_counts = _oracle_rdf_frames(z["coords"][ev1.indices], z["box"], nbins=50, return_counts=True)
    assert np.array_equal(ev1.avg, solo[0]) and np.array_equal(ev1.counts, solo_counts)
    # (2b) SFiso / SF2d: contiguous frame blocks, rows assembled by one all-reduce; with nav=1 one
    # rank has no frame at all and must neither raise nor leave the other waiting
    evaluate.sfiso_frames = _oracle_sfiso_frames
    evaluate.sf2d_frames = _oracle_sf2d_frames
    small = amep_b200.TensorTrajectory(z["coords"][:, :200], z["box"])
    for nav in (3, 1):
        es = evaluate.SFiso(small, skip=0.0, nav=nav, qmax=2.0, num=4)
        rows = [_oracle_sfiso_frames(z["coords"][i:i + 1, :200], z["box"], qmax=2.0, num=4) for i in es.indices]
        want = np.array([(r[0][0], r[1][0]) for r in rows])
        if nav == 1:    # fewer frames than ranks: the particles of the frame are split, sums all-reduced
            assert np.allclose(es.frames, want, rtol=1e-12, atol=1e-12) and np.array_equal(es.frames[:, 1], want[:, 1])
        else:
            assert np.array_equal(es.frames, want) and np.array_equal(es.avg, np.mean(want, axis=0)[0])
        e2 = evaluate.SF2d(small, skip=0.0, nav=nav, qmax=1.0, rotate=False)
        rows = [_oracle_sf2d_frames(z["coords"][i:i + 1, :200], z["box"], qmax=1.0) for i in e2.indices]
        want = np.array([(r[0][0], r[1][0], r[2][0]) for r in rows])
        mean = np.mean(want, axis=0)
        assert np.allclose(e2.avg, mean[0], rtol=1e-13, atol=0)          # partial sums per rank: last-bit freedom
        assert np.array_equal(e2.qx, mean[1]) and np.array_equal(e2.qy, mean[2])
        assert np.array_equal(e2.frames, want)                            # gathered on demand, exact
    # (3) evaluate.PCF2d: frames sharded, one all-reduce, average as average_func does
    evaluate.pcf2d_frames = _oracle_pcf2d_frames
    rng = np.random.default_rng(77)
    pc = np.zeros((5, 150, 3), dtype=np.float32)
    pc[:, :, :2] = rng.uniform(-6, 6, (5, 150, 2)).astype(np.float32)
    pc[:, :, 2] = 0.25
    pbox = np.array([[-6, 6], [-6, 6], [-0.5, 0.5]], dtype=np.float32)
    psi = rng.normal(size=(5, 2))
    ptraj = amep_b200.TensorTrajectory(pc.copy(), pbox)
    evp = evaluate.PCF2d(ptraj, skip=0.0, nav=4, psi=psi, nxbins=12, nybins=12, rmax=3.0)
    want = [orc_pcf2d(pc[i], pbox, psi[i]) for i in evp.indices]
    assert np.array_equal(evp.avg, np.mean(np.array([w[0] for w in want]), axis=0))
    assert np.array_equal(evp.x, want[0][1]) and np.array_equal(evp.y, want[0][2])
    assert evp.frames.shape == (len(evp.indices), 3, 12, 12)
    # (4) the per-frame pair observables (PCFangle, SpatialVelCor): frames sharded in blocks, rows assembled
    import amep_oracle as orc
    def _oracle_pcf_angle(c, box, other_coords=None, psi=None, e=None, shard=None, reduce=None, **kw):
        """Stand-in for spatialcor.pcf_angle.  With ``shard`` / ``reduce`` (fewer frames than ranks) the
        integer grids go through the all-reduce before the normalisation (rank 0 contributes the counts of
        the frame, the others zeros: the slicing of the queries itself is tested on the GPU)."""
        c = np.asarray(c)
        if shard is None:
            return orc.pcf_angle(c, box, other_coords, psi=psi, e=e, **kw)
        hist, dbins, abins = orc.pcf_angle_counts(c, box, other_coords, psi=psi, e=e, **kw)
        hist = reduce(hist if shard[0] == 0 else np.zeros_like(hist))
        bl = np.asarray(box)[:, 1] - np.asarray(box)[:, 0]
        n_other = len(c) if other_coords is None else len(other_coords)
        area = 0.5 * (dbins[1:] ** 2 - dbins[:-1] ** 2)[:, None] * (abins[1:] - abins[:-1])
        res = hist.astype(float) / area / (len(c) / (bl[0] * bl[1])) / n_other
        R, T = np.meshgrid(orc.bin_centres(dbins), orc.bin_centres(abins))
        return res.T, R, T
    evaluate.pcf_angle = _oracle_pcf_angle
    def _oracle_spatialcor(c, box, v, other_coords=None, other_values=None, shard=None, reduce=None, **kw):
        """Stand-in for spatialcor.spatialcor; with ``shard`` / ``reduce`` (fewer frames than ranks) this
        rank's slice of the second set, pair counts and sums added over the ranks before the normalisation."""
        c, v = np.asarray(c), np.asarray(v)
        if shard is None:
            return orc.spatialcor(c, box, v, other_coords, other_values, **kw)
        oc = c if other_coords is None else np.asarray(other_coords)
        ov = v if other_values is None else np.asarray(other_values)
        lo, hi = adist.shard_bounds(len(oc), shard[0], shard[1])
        cor, norm, edges = orc.spatialcor_sums(c, box, v, oc[lo:hi], ov[lo:hi], **kw)
        cor, norm = reduce(cor), reduce(norm).astype(float)
        norm[(norm == 0) & (cor == 0)] = 1
        cor = cor / norm
        return cor / cor[0], orc.bin_centres(edges)
    evaluate.spatialcor = _oracle_spatialcor
    flat = pc.copy()
    flat[:, :, 2] = 0.0
    vel = rng.normal(size=flat.shape).astype(np.float32)
    vtraj = amep_b200.TensorTrajectory(flat, pbox, velocities=vel)
    for nav in (5, 1):
        eva = evaluate.PCFangle(vtraj, skip=0.0, nav=nav, mode="x", ndbins=10, nabins=8, rmax=3.0)
        want = np.array([np.array(orc.pcf_angle(flat[i].copy(), pbox, None, ndbins=10, nabins=8, rmax=3.0)) for i in eva.indices])
        assert np.array_equal(eva.frames, want) and np.array_equal(eva.avg, np.mean(want, axis=0)[0])
        evv = evaluate.SpatialVelCor(vtraj, skip=0.0, nav=nav, rmax=4.0)
        want = np.array([np.array(orc.spatialcor(flat[i], pbox, vel[i], flat[i], vel[i], rmax=4.0)) for i in evv.indices])
        if nav == 1:    # the frame's query particles are split over the ranks: float64 partial sums
            assert np.allclose(evv.frames, want, rtol=1e-12, atol=1e-14) and evv.r.shape == want[0][1].shape
        else:
            assert np.array_equal(evv.frames, want) and evv.r.shape == want[0][1].shape
    np.save(os.path.join(tmpdir, f"ok{rank}.npy"), np.array([1]))
    dist.destroy_process_group()


def test_gloo_world2_sharding_and_allreduce(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok0.npy") and os.path.exists(tmp_path / "ok1.npy")
