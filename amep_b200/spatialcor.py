"""Drop-in for the RDF / S(q) functions of ``amep.spatialcor`` (AMEP 1.3.0).

Same signatures, argument meaning and error behaviour as
``amep.spatialcor.rdf`` (amep/spatialcor.py:421-649), ``sfiso`` (:1451-1687),
``sf2d`` (:1722-1940) and ``pcf2d`` (:726-926).  The per-particle work runs in libamepb's sm_100a
kernels; what stays on the host is the parameter prologue (bin edges, q grid,
unit vectors) and the normalisation epilogue, written as the same NumPy
expressions as the reference because they define output dtype and rounding
(SURVEY.md Appendix A).  ``coords`` may be a NumPy array or a (CUDA) torch
tensor; results are NumPy arrays like the reference's.

``njobs``, ``chunksize`` (except where it changes the result, see ``sf2d``) and
``verbose`` are accepted and ignored: there is no process pool to size.
The ``*_frames`` functions are the batched forms the evaluate classes use.
"""
from __future__ import annotations

import logging

import numpy as np

from . import _core
from .utils import (available_cpu_count, optimal_chunksize, runningmean, sq_from_sf2d,
                    unit_vector_2D, unit_vector_3D)

_log = logging.getLogger("amep_b200.spatialcor")

__all__ = ["rdf", "sfiso", "sf2d", "pcf2d", "pcf_angle", "spatialcor", "rdf_frames", "sfiso_frames", "sf2d_frames",
           "pcf2d_frames"]


def _box_numpy(box_boundary):
    if _core._is_torch(box_boundary):
        box_boundary = box_boundary.detach().cpu().numpy()
    return np.asarray(box_boundary)


# =============================================================================
# RDF
# =============================================================================
def _rdf_edges(box_boundary, nbins, rmax):
    """Bin edges as the reference builds them (amep/spatialcor.py:548-563): the
    dtype follows ``rmax`` -- a float32 scalar when it defaults from a float32
    box -- and NumPy later compares float64 distances against these values."""
    if nbins is None:
        nbins = 500
    if rmax is None:
        rmax = max(box_boundary[:, 1] - box_boundary[:, 0]) / 2
    return np.linspace(0.0, rmax, nbins + 1), nbins


def _rdf_normalise(hist, bins, box, n, n_other, dim):
    """amep/spatialcor.py:639-649."""
    if dim == 2:
        area = np.pi * (bins[1:] ** 2 - bins[:-1] ** 2)
        density = n / (box[0] * box[1])
        res = hist / area / density / n_other
    else:  # means dim=3
        volume = 4 / 3 * np.pi * (bins[1:] ** 3 - bins[:-1] ** 3)
        density = n / np.prod(box)
        res = hist / volume / density / n_other
    return res, runningmean(bins, 2)


def rdf_frames(coords, box_boundary, other_coords=None, nbins=None, rmax=None, mode="diff",
               pbc=True, return_counts=False, reduce=None, shard=None):
    """``rdf`` for a batch of frames ``[F, N, 3]`` in one library call.

    ``box_boundary``: ``(3,2)`` or ``(F,3,2)``.  Returns ``(g [F, nbins], r [F, nbins])``
    (plus the int64 counts with ``return_counts=True``).  ``reduce`` is an
    optional callable applied to the device/host histogram before the epilogue
    (the multi-GPU all-reduce hooks in here); ``shard=(index, count)`` evaluates only that
    share of the query cells of every frame (intra-frame sharding, amepb_rdf_set_shard)."""
    if mode not in ["diff", "kdtree"]:
        raise ValueError(
            f"amep.spatialcor.rdf: Invalid mode '{mode}'. Available modes "
            "are 'diff' and 'kdtree'.")
    pc = coords if isinstance(coords, _core.Particles) else _core.Particles(coords)
    po = None
    if other_coords is not None:
        po = other_coords if isinstance(other_coords, _core.Particles) else _core.Particles(other_coords)
    nf = pc.nframes
    bb = _box_numpy(box_boundary)
    per_frame_box = bb.ndim == 3
    if per_frame_box and bb.shape[0] != nf:
        raise ValueError("box_boundary must have shape (3,2) or (F,3,2)")
    # host prologue: bin edges per frame (shared when the box is)
    if per_frame_box:
        edge_rows = [_rdf_edges(bb[f], nbins, rmax) for f in range(nf)]
        nb = edge_rows[0][1]
        edge_list = [e for e, _ in edge_rows]
        edges = np.stack([np.asarray(e, dtype=np.float64) for e in edge_list])
    else:
        e, nb = _rdf_edges(bb, nbins, rmax)
        edge_list = [e] * nf
        edges = np.asarray(e, dtype=np.float64)
    if np.any(np.diff(edges, axis=-1) < 0):
        raise ValueError("`bins` must increase monotonically, when an array")
    hist, dims = _core.rdf_histograms(pc, bb, edges, other=po, pbc=pbc, shard=shard)
    if mode == "kdtree" and pbc:
        # the reference's kdtree mode counts one minimum-image distance per pair
        # (amep/pbc.py:632-655), 'diff' every image inside the pbc_points window: the two differ
        # beyond half the shortest periodic box length
        lengths = np.broadcast_to(bb[..., 1] - bb[..., 0], (nf, 3))
        short = min(float(np.min(lengths[f, :max(2, int(dims[f]))])) for f in range(nf)) if nf else np.inf
        if float(np.max(edges)) > 0.5 * short:
            _log.warning("amep_b200.spatialcor.rdf: mode='kdtree' is evaluated like mode='diff' (every periodic "
                         "image inside the pbc_points window); beyond rmax = min(box)/2 = %g the reference's "
                         "kdtree mode (minimum image only) counts fewer pairs than this result.", 0.5 * short)
    if reduce is not None:
        hist = reduce(hist)
    if _core._is_torch(hist):
        hist = hist.cpu().numpy()
    n, n_other = pc.n, (pc.n if po is None else po.n)
    if not per_frame_box and nf > 0:
        # one box, one set of edges: the reference's per-frame expression
        # hist / shell / density / Nother is elementwise, so evaluating it on all rows at once
        # gives the same bits as the frame loop (tests/test_abi_host.py pins that)
        box = bb[:, 1] - bb[:, 0]
        g_all, r_row = _rdf_normalise(hist, e, box, n, n_other, int(dims[0]))
        for dim in np.unique(dims):   # frames of another dimensionality (rare): their own shells
            if dim != dims[0]:
                rows = np.flatnonzero(dims == dim)
                g_all[rows] = _rdf_normalise(hist[rows], e, box, n, n_other, int(dim))[0]
        r_all = np.broadcast_to(r_row, (nf, nb)).copy()
    else:
        g_rows, r_rows = [], []
        for f in range(nf):
            b = bb[f]
            box = b[:, 1] - b[:, 0]
            g, r = _rdf_normalise(hist[f], edge_list[f], box, n, n_other, int(dims[f]))
            g_rows.append(g)
            r_rows.append(r)
        g_all, r_all = np.array(g_rows), np.array(r_rows)
    if return_counts:
        return g_all, r_all, hist
    return g_all, r_all


def rdf(coords, box_boundary, other_coords=None, nbins=None, rmax=None, mode="diff",
        chunksize=None, njobs=1, pbc=True, verbose=False):
    """Radial pair-distribution function of one frame; same contract as
    ``amep.spatialcor.rdf`` (amep/spatialcor.py:421-649).

    ``mode='kdtree'`` is served by the same exact-arithmetic kernel as
    ``'diff'`` (the reference's two modes agree to rounding for
    ``rmax <= min(L)/2``; its own test compares them after ``round(3)``,
    test/test_spatialcor.py:73-102)."""
    g, r = rdf_frames(coords, box_boundary, other_coords, nbins=nbins, rmax=rmax, mode=mode, pbc=pbc)
    return g[0], r[0]


# =============================================================================
# isotropic S(q)
# =============================================================================
def _sfiso_directions(twod, num):
    """Unit vectors of sfiso(mode='fast') (amep/spatialcor.py:1640-1647)."""
    if twod:
        theta = np.linspace(0, 360, num=num, endpoint=False)
        return unit_vector_2D(theta)
    theta = np.linspace(0, 180, num=int(num / 2), endpoint=False)
    phi = np.linspace(0, 360, num=num, endpoint=False)
    t, p = np.meshgrid(theta, phi)
    return unit_vector_3D(t.flatten(), p.flatten())


def _unique_directions(unit_vectors):
    """Group directions that coincide up to sign (S_u = S_-u exactly in exact
    arithmetic; rho(-q) = conj rho(q)).  Returns (representatives, weights)."""
    reps, weights = [], []
    for u in unit_vectors:
        for i, r in enumerate(reps):
            if np.max(np.abs(u - r)) <= 5e-16 or np.max(np.abs(u + r)) <= 5e-16:
                weights[i] += 1
                break
        else:
            reps.append(u)
            weights.append(1)
    return np.array(reps), np.array(weights, dtype=np.float64)


def _sfiso_chunksize(n_other, chunksize, njobs):
    """amep/spatialcor.py:1583-1590: part of the result in mode 'std' with ``accum='f32'`` -- every
    chunk of ``other_coords`` is summed on its own in float32."""
    if njobs > available_cpu_count():
        njobs = available_cpu_count()
    if chunksize is None:
        chunksize = optimal_chunksize(n_other, 2 * n_other)
    if chunksize > n_other:
        chunksize = int(n_other / njobs)
    return chunksize


def sfiso_frames(coords, box_boundary, N=None, qmax=20.0, twod=True, mode="std",
                 other_coords=None, accuracy=0.5, num=64, reduce=None, counts=None,
                 accum="f32", chunksize=None, njobs=1):
    """``sfiso`` for a batch of frames -> ``(S [F, nq], q [F, nq])``.

    Mode 'std': ``accum='f32'`` (default) follows the reference's float32 arithmetic *and* summation
    order (row after row or NumPy's pairwise sum, per chunk of ``chunksize`` particles of
    ``other_coords``; ``chunksize`` / ``njobs`` have the reference's meaning and defaults);
    ``accum='f64'`` rounds distances and arguments like the reference but evaluates and adds the
    terms in float64 (one pass over all pairs: closer to the exact sum than the reference itself,
    SURVEY.md section 0 fact 5, and the only form that can be sharded over ranks).

    ``reduce`` / ``counts``: intra-frame sharding over ranks -- every rank passes its slice of the
    particles, ``reduce`` sums the per-direction (cos, sin) sums (mode 'fast') or the Debye sums
    of the slice of ``coords`` against all of ``other_coords`` (mode 'std') over the ranks, and
    ``counts = (N_a, N_b)`` are the particle numbers of the whole frame for the normalisation."""
    modes = ["std", "fast", "fft"]
    pc = coords if isinstance(coords, _core.Particles) else _core.Particles(coords)
    po = pc if other_coords is None else (
        other_coords if isinstance(other_coords, _core.Particles) else _core.Particles(other_coords))
    nf = pc.nframes
    bb = _box_numpy(box_boundary)
    per_frame_box = bb.ndim == 3
    if mode not in modes:
        _log.warning(f"Mode '{mode}' not available. Using mode 'fast'. Available modes are {modes}.")
        mode = "fast"
    if not twod and mode == "fft":
        _log.warning("Mode 'fft' does only work for 2d systems. Since twod is False, mode 'std' is used.")
        mode = "std"
    if mode == "fft":
        # amep/spatialcor.py:1667-1685: radial average of sf2d(mode='fft')
        # (the reference overwrites its argument N with len(other_coords), amep/spatialcor.py:1579-1581)
        S2d, Kx, Ky = sf2d_frames(pc, bb, None if po is pc else po, qmax=qmax, mode="fft",
                                  Ntot=po.n, accuracy=accuracy)
        s_rows, q_rows = [], []
        for f in range(nf):
            b = bb[f] if per_frame_box else bb
            dq = 2 * np.pi / np.max(b[:, 1] - b[:, 0])
            S, q = sq_from_sf2d(S2d[f], Kx[f], Ky[f])
            keep = (q < qmax) & (q >= dq)
            s_rows.append(S[keep])
            q_rows.append(q[keep])
        if any(len(q) != len(q_rows[0]) for q in q_rows):
            raise ValueError("sfiso_frames: the box changes the number of q values between frames")
        return np.stack(s_rows), np.stack(q_rows)
    # q grid per frame (amep/spatialcor.py:1592-1598)
    q_rows, dq_rows = [], []
    for f in range(nf if per_frame_box else 1):
        b = bb[f] if per_frame_box else bb
        rmax = np.max(b[:, 1] - b[:, 0])
        dq = 2 * np.pi / rmax
        q_rows.append(np.arange(dq, qmax, dq))
        dq_rows.append(float(dq))
    nq = len(q_rows[0])
    if any(len(q) != nq for q in q_rows):
        raise ValueError("sfiso_frames: the box changes the number of q values between frames")
    if mode == "std":
        # Debye scattering formula over all ordered pairs, no periodic images
        # (amep/spatialcor.py:1363-1399, 1617-1634)
        if nq == 0:
            return np.zeros((nf, 0)), np.zeros((nf, 0))
        qarr = np.stack(q_rows) if per_frame_box else q_rows[0]
        if accum not in ("f32", "f64"):
            raise ValueError("accum must be 'f32' (the reference's order) or 'f64'")
        if accum == "f32" and reduce is None:
            cs = _sfiso_chunksize(po.n, chunksize, njobs)
            if cs < 1:
                raise ValueError("sfiso: chunksize must be positive")
            sums = _core.sq_debye_ref(pc, None if po is pc else po, qarr, twod, cs)
        else:
            # (a sum split over ranks cannot follow the reference's order)
            sums = _core.sq_debye(pc, None if po is pc else po, qarr, twod)
        if reduce is not None:
            sums = reduce(sums)
        if _core._is_torch(sums):
            sums = sums.cpu().numpy()
        na, nb_ = (pc.n, po.n) if counts is None else counts
        s = sums / np.sqrt(na * nb_)
        q = np.stack(q_rows) if per_frame_box else np.broadcast_to(q_rows[0], (nf, nq)).copy()
        return s, q
    unit_vectors = _sfiso_directions(twod, num)
    reps, weights = _unique_directions(unit_vectors)
    if per_frame_box:
        kvec = np.stack([dq * reps for dq in dq_rows])
    else:
        kvec = dq_rows[0] * reps
    if nq == 0:
        return np.zeros((nf, 0)), np.zeros((nf, 0))
    ha = _core.sq_harmonics(pc, kvec, nq)
    hb = ha if po is pc else _core.sq_harmonics(po, kvec, nq)
    if reduce is not None:
        ha = reduce(ha)
        hb = ha if po is pc else reduce(hb)
    if _core._is_torch(ha):
        ha = ha.cpu().numpy()
        hb = ha if po is pc else hb.cpu().numpy()
    # (C_a C_b + S_a S_b) / sqrt(N_a N_b), mean over directions
    # (amep/spatialcor.py:1440-1446, 1660-1664)
    na, nb_ = (pc.n, po.n) if counts is None else counts
    s_dir = (ha[..., 0] * hb[..., 0] + ha[..., 1] * hb[..., 1]) / np.sqrt(na * nb_)
    s = np.tensordot(weights, s_dir, axes=([0], [1])) / len(unit_vectors)
    q = np.stack(q_rows) if per_frame_box else np.broadcast_to(q_rows[0], (nf, nq)).copy()
    return s, q


def sfiso(coords, box_boundary, N, qmax=20.0, twod=True, njobs=1, chunksize=None, mode="std",
          other_coords=None, accuracy=0.5, num=64, verbose=False, accum="f32"):
    """Isotropic static structure factor of one frame; contract of
    ``amep.spatialcor.sfiso`` (amep/spatialcor.py:1451-1687).  ``N`` is ignored
    exactly as in the reference (:1579-1581).  ``accum``: see ``sfiso_frames``."""
    s, q = sfiso_frames(coords, box_boundary, N, qmax=qmax, twod=twod, mode=mode,
                        other_coords=other_coords, accuracy=accuracy, num=num,
                        accum=accum, chunksize=chunksize, njobs=njobs)
    return s[0], q[0]


# =============================================================================
# 2D S(q)
# =============================================================================
def _sf2d_chunksize(n, n_other, chunksize, njobs):
    """amep/spatialcor.py:1860-1871.  Part of the result in mode 'std': every
    chunk restarts a complex64 accumulator."""
    if njobs > available_cpu_count():
        njobs = available_cpu_count()
    if chunksize is None:
        chunksize = min(optimal_chunksize(n_other, n), optimal_chunksize(n, n_other))
    if chunksize > n_other:
        chunksize = int(n_other / njobs)
    elif chunksize > n:
        chunksize = int(n / njobs)
    return chunksize



def _frame_slice(p, f):
    """Frame ``f`` of a Particles batch as a one-frame batch (no copy)."""
    return _core.Particles(p.tensor[f:f + 1] if p.on_device else p.array[f:f + 1])


def _sf2d_fft_frames(pc, po, same, bb, qmax, Ntot, accuracy):
    """``sf2d(mode='fft')`` (amep/spatialcor.py:1911-1940): histogram density estimate of the
    particles (amep/continuum.py:51-139, 300-407; no periodic images: ``hde`` is called with its
    default ``pbc=False``), 2D FFT, ``S = Re(F_a F_b*)`` with the reference's normalisation
    (amep/continuum.py:250-281), cropped to ``|q| <= qmax``.

    The integer counts come from ``amepb_density_counts`` (exact ``np.histogram2d`` semantics), the
    FFT is cuFFT through ``torch.fft`` (library code, like the reference's ``scipy.fft``), the small
    array arithmetic around it repeats the reference's NumPy expressions one by one."""
    import torch
    nf = pc.nframes
    per_frame_box = bb.ndim == 3
    dmin = 2 * np.pi / 2 / qmax / (accuracy * 10)
    dev = pc.tensor.device if pc.on_device else torch.device("cuda", _core._device_of(pc, None))
    rows = []
    counts_a = counts_b = None
    if not per_frame_box:
        xe = np.arange(bb[0, 0] - 0.5 * dmin, bb[0, 1] + dmin, dmin)
        ye = np.arange(bb[1, 0] - 0.5 * dmin, bb[1, 1] + dmin, dmin)
        counts_a = _core.density_counts(pc, bb, xe, ye, pbc=False)
        counts_b = counts_a if same else _core.density_counts(po, bb, xe, ye, pbc=False)
    for f in range(nf):
        b = bb[f] if per_frame_box else bb
        if per_frame_box:
            xe = np.arange(b[0, 0] - 0.5 * dmin, b[0, 1] + dmin, dmin)
            ye = np.arange(b[1, 0] - 0.5 * dmin, b[1, 1] + dmin, dmin)
            ca = _core.density_counts(_frame_slice(pc, f), b, xe, ye, pbc=False)[0]
            cb = ca if same else _core.density_counts(_frame_slice(po, f), b, xe, ye, pbc=False)[0]
        else:
            ca, cb = counts_a[f], counts_b[f]
        ca_t = ca if _core._is_torch(ca) else torch.from_numpy(ca).to(dev)
        cb_t = ca_t if same else (cb if _core._is_torch(cb) else torch.from_numpy(cb).to(dev))
        # density fields (amep/continuum.py:400-402).  Their sums are taken by NumPy on the host,
        # on arrays with the reference's memory layout (the transposed view of an [nx, ny]
        # histogram): int(sum*dx*dy) below truncates a float that sits within rounding of an
        # integer, so even the order of the pairwise summation has to be the reference's
        def field(c_t):
            return np.ascontiguousarray(c_t.cpu().numpy().T).astype(np.float64).T / dmin ** 2
        d_host = field(ca_t)
        do_host = d_host if same else field(cb_t)
        rx, ry = runningmean(xe, 2), runningmean(ye, 2)
        dx, dy = rx[1] - rx[0], ry[1] - ry[0]
        sum_a, sum_b = np.sum(d_host), np.sum(do_host)
        Ni, Nj = int(sum_a * dx * dy), int(sum_b * dx * dy)
        # FFT of the densities (amep/continuum.py:262-266, norm='backward')
        fa = torch.fft.fftshift(torch.fft.fft2(ca_t.to(torch.float64) / dmin ** 2))
        fb = fa if same else torch.fft.fftshift(torch.fft.fft2(cb_t.to(torch.float64) / dmin ** 2))
        S = torch.real(fa * fb.conj())
        if Ntot is not None:
            S = S / float(sum_a) / float(sum_b) * Ni * Nj / Ntot
        S = S.cpu().numpy()
        qx = np.fft.fftshift(np.fft.fftfreq(S.shape[0], d=dx)) * 2 * np.pi
        qy = np.fft.fftshift(np.fft.fftfreq(S.shape[1], d=dy)) * 2 * np.pi
        Qx, Qy = np.meshgrid(qx, qy)
        i0, j0 = np.where((Qx == 0.0) & (Qy == 0.0))
        S[i0, j0] = 0
        if np.max(Qx) > qmax or np.max(Qy) > qmax:
            # only the q values in [-qmax, qmax] (amep/spatialcor.py:1922-1935)
            dqx = Qx[0, 1] - Qx[0, 0]
            dqy = Qy[1, 0] - Qy[0, 0]
            nqx, nqy = len(np.arange(dqx, qmax, dqx)), len(np.arange(dqy, qmax, dqy))
            mask = (np.abs(Qx) <= qmax) & (np.abs(Qy) <= qmax)
            sx, sy = int(2 * nqx + 1), int(2 * nqy + 1)
            S, Qx, Qy = S[mask].reshape(sx, sy), Qx[mask].reshape(sx, sy), Qy[mask].reshape(sx, sy)
        S[S.shape[0] // 2, S.shape[0] // 2] = 0
        rows.append((S, Qx, Qy))
    if any(r[0].shape != rows[0][0].shape for r in rows):
        raise ValueError("sf2d_frames: the box changes the grid between frames")
    return (np.stack([r[0] for r in rows]), np.stack([r[1] for r in rows]), np.stack([r[2] for r in rows]))


def sf2d_frames(coords, box_boundary, other_coords=None, qmax=20.0, njobs=1, mode="std", Ntot=None,
                accuracy=0.1, chunksize=None, accum="c64", reduce=None, return_modes=False,
                compact_grids=False, verbose=False, shard=None):
    """``sf2d`` for a batch of frames -> ``(S [F, ny, nx], Qx [F, ny, nx], Qy [F, ny, nx])``.

    ``shard=(index, count)`` with ``reduce`` (a sum over the ranks): intra-frame sharding of mode
    'std' -- every rank holds the frames and computes its share of the rows of q tiles in the
    reference's order over all particles, ``reduce`` adds the shares (exact: every element of rho
    comes from one rank).

    ``accum='c64'`` reproduces the reference's complex64 running sums (order and
    chunking included); ``accum='f64'`` sums in float64 (closer to the exact
    value than the reference itself)."""
    pc = coords if isinstance(coords, _core.Particles) else _core.Particles(coords)
    same = other_coords is None
    po = pc if same else (other_coords if isinstance(other_coords, _core.Particles)
                          else _core.Particles(other_coords))
    nf = pc.nframes
    bb = _box_numpy(box_boundary)
    per_frame_box = bb.ndim == 3
    if Ntot is None:
        Ntot = pc.n
    if accuracy < 0.0 or accuracy > 1.0:
        raise ValueError(f"amep.spatialcor.sf2d: Accuracy must be between 0.0 and 1.0. Got {accuracy}.")
    if mode not in ["std", "fft"]:
        raise ValueError(f"Mode '{mode}' not available. Available modes are {['std', 'fft']}.")
    if mode == "fft":
        return _sf2d_fft_frames(pc, po, same, bb, qmax, Ntot, accuracy)
    # q axis per frame (amep/spatialcor.py:1849-1857)
    q_rows = []
    for f in range(nf if per_frame_box else 1):
        b = bb[f] if per_frame_box else bb
        box = b[:, 1] - b[:, 0]
        rmax = np.abs(np.max(box))
        dq = 2 * np.pi / rmax
        q_rows.append(np.arange(dq, qmax, dq))
    nq = len(q_rows[0])
    if any(len(q) != nq for q in q_rows):
        raise ValueError("sf2d_frames: the box changes the number of q values between frames")
    cs = _sf2d_chunksize(pc.n, po.n, chunksize, njobs)
    qpos = np.stack(q_rows) if per_frame_box else q_rows[0]
    rho = _core.sf2d_modes(pc, qpos, nq, accum=accum, chunksize=cs, shard=shard)
    rho_other = rho if same else _core.sf2d_modes(po, qpos, nq, accum=accum, chunksize=cs, shard=shard)
    if reduce is not None:
        rho = reduce(rho)
        rho_other = rho if same else reduce(rho_other)
    if _core._is_torch(rho):
        rho = rho.cpu().numpy()
        rho_other = rho if same else rho_other.cpu().numpy()
    # S = Re(rho rho_other*) / Ntot in NumPy, like the reference (:1908) -- its
    # complex64 multiply decides the last bit (SURVEY.md A.3 iii)
    s = np.real(rho * rho_other.conj()) / Ntot
    qx_rows, qy_rows = [], []
    for q in q_rows:
        axis = np.append(np.flip(-q), q)
        gx, gy = np.meshgrid(axis, axis)
        qx_rows.append(gx)
        qy_rows.append(gy)
    if per_frame_box:
        qx_all, qy_all = np.stack(qx_rows), np.stack(qy_rows)
    elif compact_grids:
        # one box: one pair of grids for all frames, ``[1, ny, nx]`` (evaluate.SF2d)
        qx_all, qy_all = qx_rows[0][None], qy_rows[0][None]
    else:
        qx_all = np.broadcast_to(qx_rows[0], (nf,) + qx_rows[0].shape).copy()
        qy_all = np.broadcast_to(qy_rows[0], (nf,) + qy_rows[0].shape).copy()
    if return_modes:
        return s, qx_all, qy_all, rho, rho_other
    return s, qx_all, qy_all


def sf2d(coords, box_boundary, other_coords=None, qmax=20.0, njobs=1, mode="std", Ntot=None,
         accuracy=0.1, chunksize=None, verbose=False, accum="c64"):
    """2D static structure factor of one frame; contract of
    ``amep.spatialcor.sf2d`` (amep/spatialcor.py:1722-1940)."""
    s, qx, qy = sf2d_frames(coords, box_boundary, other_coords, qmax=qmax, njobs=njobs, mode=mode,
                            Ntot=Ntot, accuracy=accuracy, chunksize=chunksize, accum=accum)
    return s[0], qx[0], qy[0]


# =============================================================================
# PCF2D
# =============================================================================
def _pcf2d_angle(psi):
    """Angle that orients the x axis along ``psi`` (amep/spatialcor.py:866-872)."""
    angle = 0.0
    if psi is not None:
        ex = np.array([1, 0])
        angle = np.arccos(np.dot(ex, psi) / np.sqrt(psi[0] ** 2 + psi[1] ** 2))
        if psi[1] < 0:
            angle = 2 * np.pi - angle
    return angle


def _zero_last_column(data):
    """``coords[:,-1] = 0.0`` (amep/spatialcor.py:846-848): the reference flattens the caller's
    arrays in place, and so does this."""
    if isinstance(data, _core.Particles):
        data = data.tensor if data.on_device else data.array
    data[..., -1] = 0.0


def pcf2d_frames(coords, box_boundary, other_coords=None, nxbins=None, nybins=None, rmax=None,
                 psi=None, zero_last_column=True, shard=None, reduce=None):
    """``pcf2d`` for a batch of frames ``[F, N, 3]`` in one library call.

    ``shard=(index, count)`` with ``reduce`` (a sum over the ranks): intra-frame sharding -- this rank
    pairs its slice of the query particles with all targets, ``reduce`` adds the integer grids of the
    ranks before the normalisation (exact).

    ``box_boundary``: ``(3,2)`` or ``(F,3,2)``; ``psi``: ``None``, ``(2,)`` or ``(F,2)``.  Returns
    ``(g [F, nx, ny], X [F, ny, nx], Y [F, ny, nx])``: per frame what ``amep.spatialcor.pcf2d``
    returns.  The bin edges follow each frame's box when ``rmax`` is None.  ``zero_last_column=False``
    leaves the inputs untouched (``evaluate.PCF2d`` works on views of the trajectory)."""
    same = other_coords is None
    if zero_last_column:
        # the reference's side effect on the caller's arrays; the kernel never reads that column
        _zero_last_column(coords)
        if not same:
            _zero_last_column(other_coords)
    pc = coords if isinstance(coords, _core.Particles) else _core.Particles(coords)
    po = pc if same else (other_coords if isinstance(other_coords, _core.Particles)
                          else _core.Particles(other_coords, "other_coords"))
    nf = pc.nframes
    bb_all = _box_numpy(box_boundary)
    if nxbins is None:
        nxbins = 500
    if nybins is None:
        nybins = 500
    if psi is not None:
        psi = np.asarray(psi)
    xbins_all, ybins_all, centers, rots, boxes = [], [], [], [], []
    for f in range(nf):
        bb = bb_all[f] if bb_all.ndim == 3 else bb_all
        box = bb[:, 1] - bb[:, 0]
        # divide by 2*np.sqrt(2) to allow averaging with rotation (amep/spatialcor.py:857-861;
        # the floor division is the reference's)
        r = max(box) // (2 * np.sqrt(2)) if rmax is None else rmax
        angle = _pcf2d_angle(None if psi is None else (psi[f] if psi.ndim == 2 else psi))
        xbins_all.append(np.linspace(-r, r, nxbins + 1))
        ybins_all.append(np.linspace(-r, r, nybins + 1))
        centers.append(np.mean(bb, axis=1))
        theta = -angle
        rots.append((1.0 if angle != 0.0 else 0.0, np.cos(theta), np.sin(theta)))
        boxes.append(box)
    center_f32 = nf > 0 and centers[0].dtype == np.float32
    queries, query_offset = (None if same else po), None
    if shard is not None:
        from .dist import shard_bounds
        lo, hi = shard_bounds(po.n, shard[0], shard[1])
        src = po.tensor if po.on_device else po.array
        queries = _core.Particles(src[:, lo:hi].contiguous() if po.on_device else np.ascontiguousarray(src[:, lo:hi]))
        query_offset = lo if same else None     # (a slice of the targets themselves: equal global indices are skipped)
    if shard is not None and queries.n == 0:
        counts = np.zeros((nf, nxbins, nybins), dtype=np.int64)
    else:
        counts = _core.pcf2d_counts(
            pc, queries, np.array(centers, dtype=np.float64).reshape(nf, 3), center_f32,
            np.array(rots, dtype=np.float64).reshape(nf, 3) if any(r[0] for r in rots) else None,
            np.array(xbins_all, dtype=np.float64).reshape(nf, nxbins + 1),
            np.array(ybins_all, dtype=np.float64).reshape(nf, nybins + 1), query_offset=query_offset)
    if reduce is not None:
        counts = reduce(counts)
    if _core._is_torch(counts):
        counts = counts.cpu().numpy()
    g, xs, ys = [], [], []
    for f in range(nf):
        xbins, ybins, box = xbins_all[f], ybins_all[f], boxes[f]
        hist = counts[f].astype(float)
        # normalization (amep/spatialcor.py:914-926)
        area = (xbins[1:] - xbins[:-1])[:, None] * (ybins[1:] - ybins[:-1])
        density = pc.n / (box[0] * box[1])
        g.append(hist / area / density / po.n)
        X, Y = np.meshgrid(runningmean(xbins, 2), runningmean(ybins, 2))
        xs.append(X)
        ys.append(Y)
    return np.array(g), np.array(xs), np.array(ys)


def pcf2d(coords, box_boundary, other_coords=None, nxbins=None, nybins=None, rmax=None, psi=None,
          njobs=1, pbc=True, verbose=False, chunksize=None):
    """2D pair correlation function g(x, y) of one frame; contract of ``amep.spatialcor.pcf2d``
    (amep/spatialcor.py:726-926), including what the reference actually does with its arguments:
    plain difference vectors ``other_coords[n] - coords`` without minimum image (``pbc`` is
    accepted and has no effect, amep/spatialcor.py:700-713), the last column of the caller's
    arrays set to zero in place (:846-848) and ``rmax = max(box)//(2*sqrt(2))`` by default (:861)."""
    g, X, Y = pcf2d_frames(coords, box_boundary, other_coords, nxbins=nxbins, nybins=nybins, rmax=rmax, psi=psi)
    return g[0], X[0], Y[0]


# =============================================================================
# PCF ANGLE
# =============================================================================
def pcf_angle(coords, box_boundary, other_coords=None, ndbins=500, nabins=100, rmax=None, psi=None, njobs=1,
              pbc=True, verbose=False, chunksize=None, e=np.array([1.0, 0.0, 0.0]), shard=None, reduce=None):
    """Angle dependent radial pair correlation function g(r, theta) of one 2D frame; contract of
    ``amep.spatialcor.pcf_angle`` (amep/spatialcor.py:1044-1318): minimum-image difference vectors,
    angle to the direction ``e`` (one vector or one per particle), rotation by the ``psi`` angle;
    the z column of the caller's arrays must be zero (ValueError otherwise) and is set to zero in
    place like the reference does (:1197-1203).  Returns ``(g.T [na, nd], R, T)``.

    ``e`` is evaluated in float64 (a float32 ``e``, e.g. trajectory orientations, makes the
    reference compute the angles in float32: counts then agree except for pairs within float32
    rounding of an angle edge).

    ``shard=(index, count)`` with ``reduce`` (a sum over the ranks): intra-frame sharding -- this rank
    pairs its slice of the query particles (and of a per-particle ``e``) with all targets, ``reduce``
    adds the integer grids before the normalisation (exact)."""
    same = other_coords is None

    def host_view(c):
        return c.detach().cpu().numpy() if _core._is_torch(c) else np.asarray(c)

    for c in ((coords,) if same else (coords, other_coords)):
        z = c[..., -1]
        nonzero = bool((z != 0).any()) if _core._is_torch(c) else bool(np.any(np.asarray(z) != 0.0))
        if nonzero:
            raise ValueError("amep.spatialcor.pcf_angle: Input coordinates are not 2D. "
                             "The z-coordinate must be zero for all particles.")
        c[..., -1] = 0.0
    bb = _box_numpy(box_boundary)
    box = bb[:, 1] - bb[:, 0]
    n = int(coords.shape[0])
    n_other = n if same else int(other_coords.shape[0])
    e = np.asarray(e)
    if e.ndim == 1:
        if e.shape[0] != 3:
            raise ValueError("amep.spatialcor.pcf_angle: Invalid shape of e. "
                             "If e is given as 1D array, it must have shape (3,).")
    elif e.ndim == 2:
        if e.shape != (n, 3):
            raise ValueError("amep.spatialcor.pcf_angle: Invalid shape of e. "
                             "If e is given as 2D array, it must have shape (N,3), "
                             "where N is the number of particles.")
    else:
        raise ValueError("amep.spatialcor.pcf_angle: Invalid shape of e. e must be either 1D or 2D array.")
    if ndbins is None:
        ndbins = 500
    if nabins is None:
        nabins = 100
    if rmax is None:
        rmax = max(box) / 2
    angle = _pcf2d_angle(psi)       # amep/spatialcor.py:1235-1241, the same expression as pcf2d's
    dbins = np.linspace(0.0, rmax, ndbins + 1)
    abins = np.linspace(0.0, 2 * np.pi, nabins + 1)
    if shard is None:
        hist = _core.pcf_angle_counts(coords, None if same else other_coords, same, bb, pbc, e, angle, dbins, abins)
    else:
        from .dist import shard_bounds
        if n_other < shard[1]:     # (every rank sees the same numbers: no rank is left waiting)
            raise ValueError("amep_b200.spatialcor.pcf_angle: more ranks than particles")
        lo, hi = shard_bounds(n_other, shard[0], shard[1])
        q = (coords if same else other_coords)[lo:hi]
        hist = _core.pcf_angle_counts(coords, q, False, bb, pbc, e[lo:hi] if e.ndim == 2 else e, angle, dbins, abins,
                                      query_offset=lo if same else None)
    if reduce is not None:
        hist = reduce(hist)
    hist = hist.astype(float)
    # normalization (amep/spatialcor.py:1306-1318)
    area = 0.5 * (dbins[1:] ** 2 - dbins[:-1] ** 2)[:, None] * (abins[1:] - abins[:-1])
    density = n / (box[0] * box[1])
    res = hist / area / density / n_other
    R, T = np.meshgrid(runningmean(dbins, 2), runningmean(abins, 2))
    return res.T, R, T


# =============================================================================
# GENERAL SPATIAL CORRELATION FUNCTION
# =============================================================================
def spatialcor(coords, box_boundary, values, other_coords=None, other_values=None, dbin_edges=None, ndists=None,
               chunksize=None, njobs=1, rmax=None, verbose=False, pbc=True, shard=None, reduce=None):
    """Spatial correlation function of per-particle values (scalars, vectors, complex numbers) of
    one frame; contract of ``amep.spatialcor.spatialcor`` (amep/spatialcor.py:141-342): for every
    distance bin the average of ``conj(other_value) * value`` over the pairs whose minimum-image
    distance (the reference's periodic KD-tree, distances below ``max(box)/2``) falls into it,
    normalised by the first bin.  The pair counts are exact; the sums are float64 (the reference
    sums float32 values in float32).

    ``shard=(index, count)`` with ``reduce`` (a sum over the ranks): intra-frame sharding -- this rank
    takes its slice of the second set (the query particles of the periodic tree), ``reduce`` adds pair
    counts and sums of the ranks before the normalisation (counts exact, sums to float64 rounding)."""
    bb = _box_numpy(box_boundary)
    box = bb[:, 1] - bb[:, 0]
    vshape = tuple(values.shape)
    if len(vshape) != 1 and len(vshape) != 2:
        raise ValueError(f"amep.spatialcor.spatialcor: values have an invalid shape {vshape}. The shape must be (N,) "
                         "for scalar values and (N,d) for vector quantities.")
    # amep/spatialcor.py:272-285: a missing partner defaults to the first set (the two error
    # branches that follow in the reference are unreachable)
    if other_coords is not None and other_values is None:
        other_values = values
    if other_coords is None and other_values is not None:
        other_coords = coords
    if other_coords is not None and int(other_values.shape[0]) != int(other_coords.shape[0]):
        raise ValueError("amep.spatialcor.spatialcor: other_values must hold one row per particle of other_coords")
    if rmax is None:
        rmax = max(box) / 2.
    if dbin_edges is None and ndists is not None:
        dbin_edges = np.linspace(0.0, rmax, ndists)
    elif dbin_edges is None and ndists is None:
        dbin_edges = np.arange(0.0, rmax, 1.05)
    dbin_edges = np.asarray(dbin_edges)
    if shard is not None:
        from .dist import shard_bounds
        oc = coords if other_coords is None else other_coords
        ov = values if other_values is None else other_values
        if int(oc.shape[0]) < shard[1]:     # (every rank sees the same numbers: no rank is left waiting)
            raise ValueError("amep_b200.spatialcor.spatialcor: more ranks than particles")
        lo, hi = shard_bounds(int(oc.shape[0]), shard[0], shard[1])
        other_coords, other_values = oc[lo:hi], ov[lo:hi]
    norm, cor, is_complex = _core.spatialcor_sums(coords, values, other_coords, other_values, bb, pbc, dbin_edges)
    if reduce is not None:
        norm, cor = reduce(norm), reduce(cor)
    norm = norm.astype(float)
    if is_complex:
        cor = cor.astype(np.complex64)      # the reference accumulates complex values in complex64 (:72-73)
    else:
        cor = cor.real.copy()
    norm[(norm == 0) & (cor == 0)] = 1
    cor = cor / norm
    if cor.dtype in [np.complex64, np.complex128, complex]:
        cor = np.real(cor)
    return cor / cor[0], runningmean(dbin_edges, 2)
