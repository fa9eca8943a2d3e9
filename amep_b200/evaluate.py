"""Drop-in for ``amep.evaluate.RDF`` / ``SFiso`` / ``SF2d`` / ``PCF2d`` (AMEP 1.3.0).

Same constructor arguments and result properties as amep/evaluate.py:406-609
(RDF), :1183-1522 (SF2d) and :1525-1838 (SFiso).  Where the reference maps a
per-frame function over the selected frames with ``average_func``
(amep/utils.py:80-187), these classes select the same frames, hand all of them
to the batched kernels in one call, shard them over the ranks of a
``torch.distributed`` job when there is one, and average exactly like
``np.mean(np.array(results), axis=0)``.

``max_workers``, ``njobs`` and ``chunksize`` (except sf2d's, which is part of
the result) are accepted for compatibility and have no effect.
"""
from __future__ import annotations

import numpy as np

from . import _core, dist
from .base import BaseEvaluation
from .spatialcor import pcf2d_frames, rdf_frames, sf2d_frames, sfiso_frames
from .trajectory import TensorTrajectory
from .utils import evaluated_indices

__all__ = ["RDF", "SFiso", "SF2d", "PCF2d"]


def _gather(traj, indices, ptype):
    """Coordinates ``[F, n, 3]`` and boxes of the selected frames."""
    if isinstance(traj, TensorTrajectory):
        try:
            return traj.stack(indices, ptype)
        except ValueError:
            pass
    frames = traj[np.asarray(indices)]
    coords = [f.coords(ptype=ptype) for f in frames]
    boxes = [np.asarray(f.box) for f in frames]
    if len({tuple(c.shape) for c in coords}) != 1:
        raise ValueError("the number of selected particles changes between frames; "
                         "evaluate such trajectories frame by frame with amep_b200.spatialcor")
    if _core._is_torch(coords[0]):
        import torch
        stack = torch.stack(list(coords))
    else:
        stack = np.stack([np.asarray(c) for c in coords])
    same_box = all(b.dtype == boxes[0].dtype and np.array_equal(b, boxes[0]) for b in boxes)
    return stack, (boxes[0] if same_box else np.stack(boxes))


def _local_block(nsel, distributed, group):
    rank, ws = dist.world(group)
    if distributed is False or ws == 1:
        return 0, nsel, 1
    lo, hi = dist.shard_bounds(nsel, rank, ws)
    return lo, hi, ws


def _box_rows(box, lo, hi):
    return box[lo:hi] if np.ndim(box) == 3 else box


class _FrameAverage(BaseEvaluation):
    """Common frame selection / sharding / averaging."""

    def _select(self, traj, skip, nav):
        self._indices = evaluated_indices(len(traj), skip, nav)
        return self._indices

    def _finish(self, traj, per_frame):
        """``per_frame``: list (one entry per selected frame) of tuples of arrays;
        mirrors amep/utils.py:184-186."""
        evaluated = np.array(per_frame)
        self._frames = evaluated
        self._times = np.asarray(traj.times)[self._indices]
        return np.mean(evaluated, axis=0)


class RDF(_FrameAverage):
    """Radial pair distribution function, time-averaged over frames
    (contract of amep/evaluate.py:406-609)."""

    def __init__(self, traj, skip: float = 0.0, nav: int = 10, ptype=None, other=None,
                 max_workers=1, distributed=None, group=None, **kwargs) -> None:
        super().__init__()
        self.name = "rdf"
        kwargs.pop("njobs", None)
        kwargs.pop("chunksize", None)
        kwargs.pop("verbose", None)
        idx = self._select(traj, skip, nav)
        nsel = len(idx)
        lo, hi, ws = _local_block(nsel, distributed, group)
        counts = None
        nbins = None
        if hi > lo:
            coords, box = _gather(traj, idx[lo:hi], ptype)
            other_coords = None if other is None else _gather(traj, idx[lo:hi], other)[0]
            g, r, counts = rdf_frames(coords, box, other_coords, return_counts=True, **kwargs)
            nbins = g.shape[1]
        if ws > 1:
            # one all-reduce over [nsel, 2, nbins] float64 rows: g and r of every frame
            # (rows of other ranks are zero, so the sum is exact), plus the counts
            import torch
            nb = torch.tensor([0 if nbins is None else nbins])
            if torch.distributed.get_backend(group) == "nccl":
                nb = nb.cuda()
            torch.distributed.all_reduce(nb, op=torch.distributed.ReduceOp.MAX, group=group)
            nbins = int(nb.item())
            local = np.zeros((hi - lo, 3, nbins))
            if hi > lo:
                local[:, 0], local[:, 1], local[:, 2] = g, r, counts
            table = dist.assemble_rows(local, lo, nsel, group)
            g, r, counts = table[:, 0], table[:, 1], table[:, 2].astype(np.int64)
        self._counts = counts
        res = self._finish(traj, [(g[f], r[f]) for f in range(nsel)])
        self._avg = res[0]
        self._r = res[1]

    @property
    def r(self):
        """Distances (bin centres)."""
        return self._r

    @property
    def frames(self):
        """``(g, r)`` of every evaluated frame, shape ``(nav, 2, nbins)``."""
        return self._frames

    @property
    def times(self):
        return self._times

    @property
    def avg(self):
        """Time-averaged g(r)."""
        return self._avg

    @property
    def indices(self):
        return self._indices

    @property
    def counts(self):
        """Integer pair counts per frame (extension; not in the reference)."""
        return self._counts


class SFiso(_FrameAverage):
    """Isotropic static structure factor, time-averaged (amep/evaluate.py:1525-1838)."""

    def __init__(self, traj, skip: float = 0.0, nav: int = 10, qmax: float = 20.0, twod: bool = True,
                 njobs: int = 1, chunksize: int = 1000, mode: str = "fast", ptype=None, other=None,
                 accuracy: float = 0.5, num: int = 8, ftype=None, max_workers=1,
                 distributed=None, group=None) -> None:
        super().__init__()
        self.name = "sfiso"
        if ftype is not None or type(traj).__name__ == "FieldTrajectory":
            raise TypeError("amep_b200.evaluate.SFiso: field trajectories are not on the B200 hot path "
                            "(SURVEY.md section 8f)")
        idx = self._select(traj, skip, nav)
        nsel = len(idx)
        lo, hi, ws = _local_block(nsel, distributed, group)
        s = q = None
        if hi > lo:
            coords, box = _gather(traj, idx[lo:hi], ptype)
            other_coords = None if other is None else _gather(traj, idx[lo:hi], other)[0]
            s, q = sfiso_frames(coords, box, None, qmax=qmax, twod=twod, mode=mode,
                                other_coords=other_coords, accuracy=accuracy, num=num)
        if ws > 1:
            s, q = _assemble_pair(s, q, lo, hi, nsel, group)
        res = self._finish(traj, [(s[f], q[f]) for f in range(nsel)])
        self._avg = res[0]
        self._q = res[1]

    @property
    def q(self):
        return self._q

    @property
    def frames(self):
        return self._frames

    @property
    def times(self):
        return self._times

    @property
    def avg(self):
        return self._avg

    @property
    def indices(self):
        return self._indices


def _assemble_pair(a, b, lo, hi, nsel, group):
    """All-reduce two equally shaped per-frame float arrays as one table."""
    import torch
    shape = torch.tensor(list(a.shape[1:]) if a is not None else [0] * 8)[:8]
    pad = torch.zeros(8, dtype=torch.int64)
    pad[:len(shape)] = shape
    ndim = torch.tensor([0 if a is None else a.ndim - 1])
    if torch.distributed.get_backend(group) == "nccl":
        pad, ndim = pad.cuda(), ndim.cuda()
    torch.distributed.all_reduce(pad, op=torch.distributed.ReduceOp.MAX, group=group)
    torch.distributed.all_reduce(ndim, op=torch.distributed.ReduceOp.MAX, group=group)
    tail = tuple(int(x) for x in pad[:int(ndim.item())].tolist())
    local = np.zeros((hi - lo, 2) + tail)
    if hi > lo:
        local[:, 0], local[:, 1] = a, b
    table = dist.assemble_rows(local, lo, nsel, group)
    return table[:, 0], table[:, 1]


class SF2d(_FrameAverage):
    """Two-dimensional static structure factor, time-averaged (amep/evaluate.py:1183-1522).

    ``rotate`` defaults to True as in the reference (a psi_6 alignment prologue,
    amep/evaluate.py:1349-1377).  That prologue is a next row of the scope table
     and not built yet, so the default *raises* instead of silently
    computing the unrotated structure factor: pass ``rotate=False`` explicitly."""

    def __init__(self, traj, skip: float = 0.0, nav: int = 10, ptype=None, other=None,
                 rotate: bool = True, ftype=None, max_workers=1, distributed=None, group=None,
                 **kwargs) -> None:
        super().__init__()
        self.name = "sf2d"
        if ftype is not None or type(traj).__name__ == "FieldTrajectory":
            raise TypeError("amep_b200.evaluate.SF2d: field trajectories are not on the B200 hot path "
                            "(SURVEY.md section 8f)")
        if rotate:
            raise NotImplementedError(
                "amep_b200.evaluate.SF2d: rotate=True (psi_6 alignment, amep/evaluate.py:1349-1377) "
                "is a next row of the scope table; pass rotate=False")
        kwargs.pop("verbose", None)
        idx = self._select(traj, skip, nav)
        nsel = len(idx)
        lo, hi, ws = _local_block(nsel, distributed, group)
        s = qx = qy = None
        if hi > lo:
            coords, box = _gather(traj, idx[lo:hi], ptype)
            # the reference always passes other_coords, so rho is evaluated twice
            # (amep/evaluate.py:1380-1394); identical arrays give identical sums,
            # so one evaluation serves both when ptype == other
            other_coords = None if other == ptype else _gather(traj, idx[lo:hi], other)[0]
            s, qx, qy = sf2d_frames(coords, box, other_coords, **kwargs)
            s = s.astype(np.float64)  # np.array((S, qx, qy)) upcasts the float32 S
        if ws > 1:
            import torch
            local = np.zeros((hi - lo, 3) + (() if s is None else s.shape[1:]))
            if s is None:
                raise ValueError("SF2d: more ranks than selected frames")
            local[:, 0], local[:, 1], local[:, 2] = s, qx, qy
            table = dist.assemble_rows(local, lo, nsel, group)
            s, qx, qy = table[:, 0], table[:, 1], table[:, 2]
        res = self._finish(traj, [(s[f], qx[f], qy[f]) for f in range(nsel)])
        self._avg = res[0]
        self._qx = res[1]
        self._qy = res[2]

    @property
    def qx(self):
        return self._qx

    @property
    def qy(self):
        return self._qy

    @property
    def frames(self):
        return self._frames

    @property
    def times(self):
        return self._times

    @property
    def avg(self):
        return self._avg

    @property
    def indices(self):
        return self._indices


class PCF2d(_FrameAverage):
    """Two-dimensional pair correlation function g(x, y), time-averaged (amep/evaluate.py:612-752).

    The reference orients every frame along its mean hexagonal order parameter,
    ``psi = np.mean(psi_k(frame.coords(), frame.box, k=6))`` (amep/evaluate.py:730-735), before it
    calls ``pcf2d``.  That psi_6 prologue is the same next row that ``SF2d(rotate=True)`` waits
    for, so the orientation is an argument here: ``psi`` is an array
    ``[len(traj), 2]`` of (Re psi_6, Im psi_6) per trajectory frame, or a callable
    ``frame -> (re, im)``.  Without it the class *raises* instead of silently returning the
    unrotated function; ``psi=False`` asks for exactly that (no rotation).  Everything else --
    frame selection, ``ptype`` / ``other``, the keyword arguments of ``pcf2d``, the average -- is
    the reference's."""

    def __init__(self, traj, skip: float = 0.0, nav: int = 10, ptype=None, other=None,
                 max_workers=1, psi=None, distributed=None, group=None, **kwargs) -> None:
        super().__init__()
        self.name = "pcf2d"
        if psi is None:
            raise NotImplementedError(
                "amep_b200.evaluate.PCF2d: the psi_6 orientation prologue (amep/evaluate.py:730-735) is a "
                "next row of the scope table; pass psi=[len(traj), 2] (or a callable frame -> (re, im)), "
                "or psi=False for no rotation")
        for ignored in ("verbose", "njobs", "chunksize", "pbc"):
            kwargs.pop(ignored, None)
        idx = self._select(traj, skip, nav)
        nsel = len(idx)
        lo, hi, ws = _local_block(nsel, distributed, group)
        g = X = Y = None
        if hi > lo:
            coords, box = _gather(traj, idx[lo:hi], ptype)
            other_coords = None if other is None else _gather(traj, idx[lo:hi], other)[0]
            if psi is False:
                psi_rows = None
            elif callable(psi):
                psi_rows = np.array([np.asarray(psi(f), dtype=np.float64) for f in traj[np.asarray(idx[lo:hi])]])
            else:
                psi_rows = np.asarray(psi, dtype=np.float64)[np.asarray(idx[lo:hi])]
            # pcf2d zeroes the last column of what it is given in place (amep/spatialcor.py:846-848);
            # in the reference that hits the fresh array frame.coords() returns, here the stacks
            # may be views of the trajectory, so the side effect is switched off
            g, X, Y = pcf2d_frames(coords, box, other_coords, psi=psi_rows, zero_last_column=False, **kwargs)
        if ws > 1:
            if g is None:
                raise ValueError("PCF2d: more ranks than selected frames")
            if g.shape[1:] != X.shape[1:]:
                raise ValueError("PCF2d: nxbins and nybins must be equal (the reference's np.array(results) "
                                 "needs g and the bin-centre grids to have one shape)")
            local = np.zeros((hi - lo, 3) + g.shape[1:])
            local[:, 0], local[:, 1], local[:, 2] = g, X, Y
            table = dist.assemble_rows(local, lo, nsel, group)
            g, X, Y = table[:, 0], table[:, 1], table[:, 2]
        res = self._finish(traj, [(g[f], X[f], Y[f]) for f in range(nsel)])
        self._avg = res[0]
        self._x = res[1]
        self._y = res[2]

    @property
    def x(self):
        return self._x

    @property
    def y(self):
        return self._y

    @property
    def frames(self):
        return self._frames

    @property
    def times(self):
        return self._times

    @property
    def avg(self):
        return self._avg

    @property
    def indices(self):
        return self._indices
