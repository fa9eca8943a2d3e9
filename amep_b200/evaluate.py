"""Drop-in for ``amep.evaluate.RDF`` / ``SFiso`` / ``SF2d`` / ``PCF2d`` (AMEP 1.3.0).

Same constructor arguments and result properties as amep/evaluate.py:406-609
(RDF), :612-752 (PCF2d), :1183-1522 (SF2d) and :1525-1838 (SFiso).  Where the
reference maps a per-frame function over the selected frames with ``average_func``
(amep/utils.py:80-187), these classes select the same frames, hand them to the
batched kernels in a few large calls, shard them over the ranks of a
``torch.distributed`` job when there is one, and average exactly like
``np.mean(np.array(results), axis=0)``.

Data path.  A ``TensorTrajectory`` (arrays already in host memory or HBM) is read in
place.  Any other trajectory object (the reference's ``ParticleTrajectory``: one HDF5
read per ``frame.coords()`` call, amep/base.py:941-979) is *streamed*: a producer thread
pulls the next block of frames into a pinned staging buffer while the GPU works on the
current one (``_streamed_batches``), so wall time is max(read, compute) rather than the
sum and host memory stays at a few staging buffers instead of the whole selection.

Multi-GPU.  Frames are the unit of sharding (contiguous blocks per rank, one
all-reduce).  With fewer selected frames than ranks the work is sharded *inside* the
frame instead, every rank holding the frame: the RDF by query cell runs
(``amepb_rdf_set_shard``; integer histograms, exact) -- the split the reference makes over
query chunks (amep/spatialcor.py:587-600) --, SF2d by rows of q tiles (``amepb_sf2d_set_shard``,
exact), SFiso by particles (sums to float64 rounding), PCF2d / PCFangle / SpatialVelCor /
HexOrderCor by query particles (integer grids and pair counts exact).

``max_workers`` and ``njobs`` are accepted for compatibility and have no effect; ``chunksize``
matters where it is part of the reference's result (sf2d's complex64 and sfiso('std')'s float32
partial sums).
"""
from __future__ import annotations

import queue
import threading

import numpy as np

from . import _core, dist
from .base import BaseEvaluation
from .order import psi6_mean, psi6_orientation, rotated_periodic_crop
from .order import psi_k
from .spatialcor import pcf2d_frames, pcf_angle, rdf_frames, sf2d_frames, sfiso_frames, spatialcor
from .trajectory import TensorTrajectory
from .utils import evaluated_indices

__all__ = ["RDF", "SFiso", "SF2d", "PCF2d", "PCFangle", "SpatialVelCor", "HexOrderCor"]

# bytes of coordinates per streamed block (per species); three staging buffers are in flight
STREAM_BLOCK_BYTES = 256 << 20


# =============================================================================
# frame access
# =============================================================================
def _same_box(boxes):
    first = boxes[0]
    if all(b.dtype == first.dtype and np.array_equal(b, first) for b in boxes):
        return first
    return np.stack(boxes)


def _pinned_empty(shape, dtype):
    """Host staging buffer, page-locked when a CUDA device is there (the library's copies from it
    then run asynchronously at full link speed)."""
    try:
        import torch
        if torch.cuda.is_available():
            tdt = {np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64}[np.dtype(dtype)]
            return torch.empty(tuple(shape), dtype=tdt).pin_memory().numpy()
    except Exception:  # pragma: no cover - no torch / no pinned memory: plain pages still work
        pass
    return np.empty(shape, dtype=dtype)


def _streamed_batches(traj, indices, ptype, other, want_other, block_bytes=None):
    """Blocks of frames of a generic trajectory: ``(coords [f,n,3], other or None, box)``.

    The reference reads frame after frame inside its serial loop (amep/utils.py:156-158, one
    ``h5py.File`` open + decompression per ``frame.coords()``, amep/base.py:941-979).  Here a
    producer thread fills the next staging buffer while the caller computes on the current one."""
    frames = traj[np.asarray(indices)]
    if len(frames) == 0:
        return
    first = frames[0].coords(ptype=ptype)
    if _core._is_torch(first):
        # frames that are tensors already (possibly in HBM): nothing to stream
        import torch
        coords = torch.stack([f.coords(ptype=ptype) for f in frames])
        oc = torch.stack([f.coords(ptype=other) for f in frames]) if want_other else None
        yield coords, oc, _same_box([np.asarray(f.box) for f in frames])
        return
    first = np.asarray(first)
    if first.dtype not in (np.float32, np.float64):
        first = first.astype(np.float64)
    shape_c, dtype_c = first.shape, first.dtype
    shape_o = dtype_o = None
    per_frame = first.nbytes
    if want_other:
        fo = np.asarray(frames[0].coords(ptype=other))
        if fo.dtype not in (np.float32, np.float64):
            fo = fo.astype(np.float64)
        shape_o, dtype_o = fo.shape, fo.dtype
        per_frame = max(per_frame, fo.nbytes)
    block = int(max(1, min(len(frames), (block_bytes or STREAM_BLOCK_BYTES) // max(1, per_frame))))
    nbuf = 3 if len(frames) > 2 * block else (2 if len(frames) > block else 1)
    bufs_c = [_pinned_empty((block,) + shape_c, dtype_c) for _ in range(nbuf)]
    bufs_o = [_pinned_empty((block,) + shape_o, dtype_o) for _ in range(nbuf)] if want_other else None
    free, ready = queue.Queue(), queue.Queue(maxsize=nbuf)
    for b in range(nbuf):
        free.put(b)
    stop = threading.Event()

    def producer():
        try:
            for lo in range(0, len(frames), block):
                b = free.get()
                if stop.is_set():
                    return
                hi = min(len(frames), lo + block)
                boxes = []
                for k, fr in enumerate(frames[lo:hi]):
                    c = np.asarray(fr.coords(ptype=ptype))
                    if c.shape != shape_c:
                        raise ValueError("the number of selected particles changes between frames; "
                                         "evaluate such trajectories frame by frame with amep_b200.spatialcor")
                    bufs_c[b][k] = c
                    if want_other:
                        o = np.asarray(fr.coords(ptype=other))
                        if o.shape != shape_o:
                            raise ValueError("the number of selected particles changes between frames; "
                                             "evaluate such trajectories frame by frame with amep_b200.spatialcor")
                        bufs_o[b][k] = o
                    boxes.append(np.asarray(fr.box))
                ready.put((b, hi - lo, boxes))
            ready.put(None)
        except BaseException as exc:  # noqa: BLE001 - handed to the consumer
            ready.put(exc)

    thread = threading.Thread(target=producer, name="amep_b200-frame-reader", daemon=True)
    thread.start()
    try:
        while True:
            item = ready.get()
            if item is None:
                break
            if isinstance(item, BaseException):
                raise item
            b, nf, boxes = item
            yield bufs_c[b][:nf], (bufs_o[b][:nf] if want_other else None), _same_box(boxes)
            free.put(b)   # the consumer has finished with this buffer (library calls are complete on return)
    finally:
        stop.set()
        free.put(0)       # unblock a producer waiting for a buffer
        thread.join(timeout=10)


def _frame_batches(traj, indices, ptype, other, want_other):
    """``(coords, other_coords or None, box)`` blocks covering ``indices`` in order."""
    if len(indices) == 0:
        return
    if isinstance(traj, TensorTrajectory):
        try:
            coords, box = traj.stack(indices, ptype)
            oc = traj.stack(indices, other)[0] if want_other else None
            yield coords, oc, box
            return
        except ValueError:
            pass
    yield from _streamed_batches(traj, indices, ptype, other, want_other)


def _gather(traj, indices, ptype):
    """Coordinates ``[F, n, 3]`` and boxes of the selected frames in one piece (copies: the blocks of
    a streamed trajectory live in staging buffers that are reused)."""
    parts = []
    for coords, _, box in _frame_batches(traj, indices, ptype, None, False):
        if isinstance(traj, TensorTrajectory) or _core._is_torch(coords):
            parts.append((coords, box))
        else:
            parts.append((np.array(coords), box))
    if len(parts) == 1:
        return parts[0]
    if _core._is_torch(parts[0][0]):
        import torch
        stack = torch.cat([p[0] for p in parts])
    else:
        stack = np.concatenate([p[0] for p in parts])
    boxes = [np.broadcast_to(p[1], (len(p[0]), 3, 2)) for p in parts]
    return stack, np.concatenate(boxes)


# =============================================================================
# sharding
# =============================================================================
class _Shard:
    """What this rank evaluates: frames ``[lo, hi)`` of the selection, or -- ``intra`` -- all of
    them with a 1/ws share of the work inside every frame."""

    def __init__(self, nsel, distributed, group, allow_intra=False):
        self.rank, self.ws = dist.world(group)
        self.group = group
        if distributed is False:
            self.rank, self.ws = 0, 1
        self.nsel = nsel
        self.intra = allow_intra and self.ws > 1 and nsel < self.ws
        if self.ws == 1 or self.intra:
            self.lo, self.hi = 0, nsel
        else:
            self.lo, self.hi = dist.shard_bounds(nsel, self.rank, self.ws)

    def all_sum(self, array):
        """Sum a NumPy array / tensor over the ranks (identity for one rank) -> NumPy."""
        if self.ws == 1:
            return array.cpu().numpy() if _core._is_torch(array) else np.asarray(array)
        return dist.all_reduce_sum(array, self.group)

    def tail_shape(self, local):
        """Shape of one row, agreed between ranks some of which may hold no frame."""
        return dist.agree_shape(None if local is None else tuple(local.shape[1:]), self.group)

    def assemble(self, local, dtype=np.float64):
        """Rows of all ranks -> ``[nsel, ...]`` (zero rows + one all-reduce; exact)."""
        if self.ws == 1 or self.intra:
            return local
        tail = self.tail_shape(local)
        rows = np.zeros((self.hi - self.lo,) + tail, dtype=dtype)
        if local is not None and self.hi > self.lo:
            rows[...] = local
        return dist.assemble_rows(rows, self.lo, self.nsel, self.group)


def _mean_of_copies(v, n):
    """``np.mean`` over ``n`` identical rows ``v`` as average_func computes it (amep/utils.py:186):
    a sequential float64 sum along the frame axis, divided by ``n`` -- not always ``v`` itself."""
    v = np.asarray(v, dtype=np.float64)
    acc = v.copy()
    for _ in range(n - 1):
        acc = acc + v
    return acc / n


class _FrameAverage(BaseEvaluation):
    """Common frame selection / averaging."""

    def _select(self, traj, skip, nav):
        self._indices = evaluated_indices(len(traj), skip, nav)
        self._times = np.asarray(traj.times)[self._indices]
        return self._indices

    @property
    def times(self):
        return self._times

    @property
    def indices(self):
        return self._indices


# =============================================================================
# RDF
# =============================================================================
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
        sh = _Shard(len(idx), distributed, group, allow_intra=True)
        rows = []   # (g, r, counts) per block
        extra = {}
        if sh.intra:
            # fewer frames than ranks: every rank evaluates its share of the query cells of every
            # frame; the integer histograms are summed before the normalisation
            extra = dict(shard=(sh.rank, sh.ws), reduce=sh.all_sum)
        for coords, other_coords, box in _frame_batches(traj, idx[sh.lo:sh.hi], ptype, other, other is not None):
            rows.append(rdf_frames(coords, box, other_coords, return_counts=True, **extra, **kwargs))
        g = r = counts = None
        if rows:
            g, r, counts = rows[0] if len(rows) == 1 else (np.concatenate([p[k] for p in rows]) for k in range(3))
        if sh.ws == 1 or sh.intra:
            # np.array(results) of the reference: (nav, 2, nbins)
            self._frames = np.empty((len(g), 2, g.shape[1]))
            self._frames[:, 0], self._frames[:, 1] = g, r
            self._counts = counts
        else:
            local = None
            if rows:
                local = np.stack([g, r, counts.astype(np.float64)], axis=1)   # counts < 2^53: exact
            table = sh.assemble(local)
            self._frames = np.ascontiguousarray(table[:, :2])
            self._counts = table[:, 2].astype(np.int64)
        avg = np.mean(self._frames, axis=0)   # amep/utils.py:186
        self._avg = avg[0]
        self._r = avg[1]

    @property
    def r(self):
        """Distances (bin centres)."""
        return self._r

    @property
    def frames(self):
        """``(g, r)`` of every evaluated frame, shape ``(nav, 2, nbins)``."""
        return self._frames

    @property
    def avg(self):
        """Time-averaged g(r)."""
        return self._avg

    @property
    def counts(self):
        """Integer pair counts per frame (extension; not in the reference)."""
        return self._counts


# =============================================================================
# isotropic S(q)
# =============================================================================
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
        # fewer frames than ranks: the particle sums of 'fast' (and the Debye sums of 'std') are linear
        # in the particles, so every rank takes a slice of them and the sums are all-reduced
        sh = _Shard(len(idx), distributed, group, allow_intra=mode in ("fast", "std"))
        rows = []
        for coords, other_coords, box in _frame_batches(traj, idx[sh.lo:sh.hi], ptype, other, other is not None):
            extra = {}
            if sh.intra:
                na = int(coords.shape[1])
                nb_ = na if other_coords is None else int(other_coords.shape[1])
                if min(na, nb_) < sh.ws:          # (every rank sees the same numbers: no rank is left waiting)
                    raise ValueError("SFiso: more ranks than particles")
                lo_a, hi_a = dist.shard_bounds(na, sh.rank, sh.ws)
                if mode == "fast" and other_coords is not None:
                    lo_b, hi_b = dist.shard_bounds(nb_, sh.rank, sh.ws)
                    other_coords = other_coords[:, lo_b:hi_b]
                elif mode == "std" and other_coords is None:
                    other_coords = coords               # the slice pairs with all particles
                coords = coords[:, lo_a:hi_a]
                extra = dict(reduce=sh.all_sum, counts=(na, nb_))
            # (mode 'std': the reference's float32 sums per chunk of `chunksize` particles,
            # amep/evaluate.py:1699-1724 -- unless the frame is split over the ranks)
            rows.append(sfiso_frames(coords, box, None, qmax=qmax, twod=twod, mode=mode,
                                     other_coords=other_coords, accuracy=accuracy, num=num,
                                     chunksize=chunksize, njobs=njobs, **extra))
        local = None
        if rows:
            s, q = (np.concatenate([p[k] for p in rows]) for k in range(2))
            local = np.stack([s, q], axis=1)
        self._frames = sh.assemble(local)
        avg = np.mean(self._frames, axis=0)
        self._avg = avg[0]
        self._q = avg[1]

    @property
    def q(self):
        return self._q

    @property
    def frames(self):
        return self._frames

    @property
    def avg(self):
        return self._avg


# =============================================================================
# 2D S(q)
# =============================================================================
class SF2d(_FrameAverage):
    """Two-dimensional static structure factor, time-averaged (amep/evaluate.py:1183-1522).

    ``rotate=True`` (the reference's default) aligns every frame with its mean hexagonal order
    parameter first (amep/evaluate.py:1349-1377): psi_6 from the six nearest neighbours
    (``amep_b200.order.psi6_orientation``, sm_100a kernel), rotation of the periodic images about
    the box centre and crop to the box (``rotated_periodic_crop``).

    Extension: ``orientation`` = array ``[len(traj)]`` of angles replaces the psi_6 angle (the
    reference's angle carries the float32 noise of its own psi_6, and S(q) at large q is sensitive
    to it: injecting the reference's angles is how the rotated path is compared to 1e-6).

    Memory: per-frame results are kept as float32 ``S`` rows (what ``sf2d`` returns) plus one pair
    of q grids; the reference's ``(nav, 3, ny, nx)`` float64 array is only materialised when
    ``frames`` is read.  Under ``torch.distributed`` only the frame *sum* is all-reduced for
    ``avg``; ``frames`` gathers the rows on demand (a collective: read it on every rank)."""

    def __init__(self, traj, skip: float = 0.0, nav: int = 10, ptype=None, other=None,
                 rotate: bool = True, ftype=None, max_workers=1, distributed=None, group=None,
                 orientation=None, **kwargs) -> None:
        super().__init__()
        self.name = "sf2d"
        if ftype is not None or type(traj).__name__ == "FieldTrajectory":
            raise TypeError("amep_b200.evaluate.SF2d: field trajectories are not on the B200 hot path "
                            "(SURVEY.md section 8f)")
        kwargs.pop("verbose", None)
        idx = self._select(traj, skip, nav)
        # fewer frames than ranks (mode 'std'): every rank takes all selected frames and a share of the rows
        # of q tiles of each (amepb_sf2d_set_shard); the rho arrays are summed over the ranks -- exact
        sh = self._shard = _Shard(len(idx), distributed, group, allow_intra=kwargs.get("mode", "std") == "std")
        if sh.intra:
            kwargs = dict(kwargs, shard=(sh.rank, sh.ws), reduce=sh.all_sum)
        multi = sh.ws > 1 and not sh.intra   # frames spread over the ranks
        s_rows, qx_rows, qy_rows, thetas = [], [], [], []
        done = 0
        # the reference always passes other_coords, so rho is evaluated twice
        # (amep/evaluate.py:1380-1394); identical arrays give identical sums,
        # so one evaluation serves both when ptype == other
        want_other = other != ptype
        for coords, other_coords, box in _frame_batches(traj, idx[sh.lo:sh.hi], ptype, other, want_other):
            if rotate:
                # the crop changes the particle number from frame to frame: one library call per frame
                block = np.asarray(idx[sh.lo + done:sh.lo + done + coords.shape[0]])
                allc = coords if ptype is None else _gather(traj, block, None)[0]   # psi_6 of *all* particles (:1352-1357)
                for f in range(coords.shape[0]):
                    b = box[f] if np.ndim(box) == 3 else box
                    theta = psi6_orientation(allc[f], b) if orientation is None else float(orientation[block[f]])
                    c = rotated_periodic_crop(coords[f], b, theta)
                    o = None
                    if want_other:
                        o = rotated_periodic_crop(other_coords[f], b, theta)
                    s, qx, qy = sf2d_frames(c[None], b, None if o is None else o[None], **kwargs)
                    s_rows.append(s), qx_rows.append(qx), qy_rows.append(qy)
                    thetas.append(theta)
                done += coords.shape[0]
            else:
                s, qx, qy = sf2d_frames(coords, box, other_coords, compact_grids=True, **kwargs)
                s_rows.append(s), qx_rows.append(qx), qy_rows.append(qy)
        self._s = np.concatenate(s_rows) if s_rows else None            # [local frames, ny, nx], as returned by sf2d
        qx = np.concatenate(qx_rows) if qx_rows else None               # [local frames or 1, ny, nx]
        qy = np.concatenate(qy_rows) if qy_rows else None
        if qx is not None and len(qx) > 1 and all(np.array_equal(qx[0], g) for g in qx[1:]) \
                and all(np.array_equal(qy[0], g) for g in qy[1:]):
            qx, qy = qx[:1], qy[:1]
        self._qx_rows, self._qy_rows = qx, qy
        self._frames = None
        nsel = len(idx)
        # ---- average (np.mean(np.array(results), axis=0): sequential float64 sums over the frames)
        shared_grid = qx is None or len(qx) == 1
        if multi:
            shared_grid = bool(dist.all_true(shared_grid, group))
        if not multi:
            self._avg = np.mean(self._s, axis=0, dtype=np.float64)
        else:
            tail = sh.tail_shape(self._s)
            part = np.zeros(tail) if self._s is None or len(self._s) == 0 else np.sum(self._s, axis=0, dtype=np.float64)
            self._avg = sh.all_sum(part) / nsel
        if shared_grid:
            if multi:   # ranks without frames take the grids from rank 0's block
                tail = sh.tail_shape(qx)
                gx = qx[0] if (qx is not None and sh.lo == 0 and sh.hi > 0) else np.zeros(tail)
                gy = qy[0] if (qy is not None and sh.lo == 0 and sh.hi > 0) else np.zeros(tail)
                both = sh.all_sum(np.stack([gx, gy]))
                self._qx_rows, self._qy_rows = both[0][None], both[1][None]
            # the grids are meshgrids of one axis: average the axis, then rebuild
            ax = _mean_of_copies(self._qx_rows[0][0, :], nsel)
            ay = _mean_of_copies(self._qy_rows[0][:, 0], nsel)
            self._qx, self._qy = np.meshgrid(ax, ay)
        else:
            fr = self.frames
            self._qx, self._qy = np.mean(fr[:, 1], axis=0), np.mean(fr[:, 2], axis=0)

    @property
    def qx(self):
        return self._qx

    @property
    def qy(self):
        return self._qy

    @property
    def frames(self):
        """``(S, Qx, Qy)`` of every evaluated frame, shape ``(nav, 3, ny, nx)`` float64 -- built on
        first access (12.6 GB for 2000 frames of a 512 x 512 grid; a collective under
        ``torch.distributed``)."""
        if self._frames is None:
            sh = self._shard
            s = None if self._s is None else self._s.astype(np.float64)   # np.array((S, qx, qy)) upcasts the float32 S
            s = sh.assemble(s)
            nsel = sh.nsel
            out = np.empty((nsel, 3) + s.shape[1:])
            out[:, 0] = s
            if len(self._qx_rows) == 1:
                out[:, 1], out[:, 2] = self._qx_rows[0], self._qy_rows[0]
            else:
                out[:, 1], out[:, 2] = sh.assemble(self._qx_rows), sh.assemble(self._qy_rows)
            self._frames = out
        return self._frames

    @property
    def avg(self):
        return self._avg


# =============================================================================
# PCF2d
# =============================================================================
class PCF2d(_FrameAverage):
    """Two-dimensional pair correlation function g(x, y), time-averaged (amep/evaluate.py:612-752).

    Like the reference, every frame is oriented along its mean hexagonal order parameter,
    ``psi = np.mean(psi_k(frame.coords(), frame.box, k=6))`` (amep/evaluate.py:730-735), computed by
    ``amep_b200.order.psi6_mean`` (sm_100a kernel).  Extension: ``psi`` may be given -- an array
    ``[len(traj), 2]`` of (Re psi_6, Im psi_6) per trajectory frame or a callable
    ``frame -> (re, im)`` -- and ``psi=False`` switches the rotation off."""

    def __init__(self, traj, skip: float = 0.0, nav: int = 10, ptype=None, other=None,
                 max_workers=1, psi=None, distributed=None, group=None, **kwargs) -> None:
        super().__init__()
        self.name = "pcf2d"
        for ignored in ("verbose", "njobs", "chunksize", "pbc"):
            kwargs.pop(ignored, None)
        idx = self._select(traj, skip, nav)
        # fewer frames than ranks: every rank takes all selected frames and a slice of the query particles
        # of each (amepb_pcf2d_set_query_offset); the integer grids are summed before the normalisation
        sh = _Shard(len(idx), distributed, group, allow_intra=True)
        if sh.intra:
            kwargs = dict(kwargs, shard=(sh.rank, sh.ws), reduce=sh.all_sum)
        rows = []
        done = 0
        for coords, other_coords, box in _frame_batches(traj, idx[sh.lo:sh.hi], ptype, other, other is not None):
            nf = coords.shape[0]
            block = np.asarray(idx[sh.lo + done:sh.lo + done + nf])
            done += nf
            if psi is False:
                psi_rows = None
            elif psi is None:
                # amep/evaluate.py:730-735: psi_6 of *all* particles of the frame
                if ptype is None:
                    allc = coords
                else:
                    allc = _gather(traj, block, None)[0]
                psi_rows = np.array([psi6_mean(allc[f], box[f] if np.ndim(box) == 3 else box)
                                     for f in range(nf)], dtype=np.float64)
            elif callable(psi):
                psi_rows = np.array([np.asarray(psi(f), dtype=np.float64) for f in traj[block]])
            else:
                psi_rows = np.asarray(psi, dtype=np.float64)[block]
            # pcf2d zeroes the last column of what it is given in place (amep/spatialcor.py:846-848);
            # in the reference that hits the fresh array frame.coords() returns, here the stacks
            # may be views of the trajectory, so the side effect is switched off
            rows.append(pcf2d_frames(coords, box, other_coords, psi=psi_rows, zero_last_column=False, **kwargs))
        local = None
        if rows:
            g, X, Y = (np.concatenate([p[k] for p in rows]) for k in range(3))
            if g.shape[1:] != X.shape[1:]:
                raise ValueError("PCF2d: nxbins and nybins must be equal (the reference's np.array(results) "
                                 "needs g and the bin-centre grids to have one shape)")
            local = np.stack([g, X, Y], axis=1)
        self._frames = sh.assemble(local)
        avg = np.mean(self._frames, axis=0)
        self._avg = avg[0]
        self._x = avg[1]
        self._y = avg[2]

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
    def avg(self):
        return self._avg


# =============================================================================
# per-frame pair observables: PCFangle, SpatialVelCor, HexOrderCor
# =============================================================================
class _PerFrame(_FrameAverage):
    """Evaluate classes whose kernels take one frame per call (all-pairs observables): the
    selected frames are sharded over the ranks in contiguous blocks, every rank loops over its
    frames, the rows are assembled with one all-reduce and averaged like average_func."""

    def _run(self, traj, skip, nav, distributed, group, compute, allow_intra=False):
        idx = self._select(traj, skip, nav)
        # allow_intra: with fewer frames than ranks every rank takes all selected
        # frames and a slice of the query particles of each; counts and sums are all-reduced before the
        # normalisation (compute(frame, shard=..., reduce=...))
        sh = _Shard(len(idx), distributed, group, allow_intra=allow_intra)
        extra = dict(shard=(sh.rank, sh.ws), reduce=sh.all_sum) if sh.intra else {}
        rows = [np.array(compute(frame, **extra)) for frame in traj[np.asarray(idx[sh.lo:sh.hi])]]
        local = np.stack(rows) if rows else None
        self._frames = sh.assemble(local)
        return np.mean(self._frames, axis=0)

    @property
    def frames(self):
        return self._frames

    @property
    def avg(self):
        return self._avg


class PCFangle(_PerFrame):
    """Angular pair correlation function g(r, theta), time-averaged (amep/evaluate.py:834-1196).
    ``mode``: ``'psi6'`` (angles relative to the mean hexagonal orientation of the frame),
    ``'orientations'`` (relative to every particle's own orientation) or ``'x'``."""

    def __init__(self, traj, skip: float = 0.0, nav: int = 10, ptype=None, other=None, mode: str = "psi6",
                 max_workers=1, distributed=None, group=None, **kwargs) -> None:
        super().__init__()
        self.name = "pcfa"
        if mode not in ["psi6", "orientations", "x"]:
            raise ValueError("Mode not recognized. Possible values are 'psi6', 'orientations', and 'x'.")
        for ignored in ("verbose", "njobs", "chunksize"):
            kwargs.pop(ignored, None)

        def compute(frame, **extra):
            # amep/evaluate.py:1046-1101
            psi = None
            e = np.array([1.0, 0.0, 0.0], dtype=float)
            if mode == "orientations":
                e = _core.to_numpy(frame.orientations(ptype=ptype if other is None else other))
            elif mode == "psi6":
                psi = psi6_mean(frame.coords(), frame.box)
            c = frame.coords(ptype=ptype)
            c = c.clone() if _core._is_torch(c) else np.array(c)       # pcf_angle zeroes z in place
            if other is None:
                return pcf_angle(c, frame.box, psi=psi, e=e, **kwargs, **extra)
            o = frame.coords(ptype=other)
            o = o.clone() if _core._is_torch(o) else np.array(o)
            return pcf_angle(c, frame.box, other_coords=o, psi=psi, e=e, **kwargs, **extra)

        res = self._run(traj, skip, nav, distributed, group, compute, allow_intra=True)
        self._avg, self._r, self._theta = res[0], res[1], res[2]

    @property
    def r(self):
        return self._r

    @property
    def theta(self):
        return self._theta


class SpatialVelCor(_PerFrame):
    """Spatial velocity correlation function, time-averaged (amep/evaluate.py:219-403)."""

    def __init__(self, traj, skip: float = 0.0, nav: int = 10, ptype=None, other=None, max_workers=1,
                 distributed=None, group=None, **kwargs) -> None:
        super().__init__()
        self.name = "svc"
        for ignored in ("verbose", "njobs", "chunksize"):
            kwargs.pop(ignored, None)

        def compute(frame, **extra):
            # amep/evaluate.py:330-338: the second set is always passed
            return spatialcor(frame.coords(ptype=ptype), frame.box, frame.velocities(ptype=ptype),
                              other_coords=frame.coords(ptype=other), other_values=frame.velocities(ptype=other),
                              **kwargs, **extra)

        res = self._run(traj, skip, nav, distributed, group, compute, allow_intra=True)
        self._avg, self._r = res[0], res[1]

    @property
    def r(self):
        return self._r


class HexOrderCor(_PerFrame):
    """Spatial correlation function of the hexagonal order parameter, time-averaged
    (amep/evaluate.py:2085-2330): psi_6 per particle (amepb_psi_k), then ``spatialcor``."""

    def __init__(self, traj, skip: float = 0.0, nav: int = 10, ptype=None, other=None, max_workers=1,
                 distributed=None, group=None, **kwargs) -> None:
        super().__init__()
        self.name = "hoc"
        for ignored in ("verbose", "njobs", "chunksize"):
            kwargs.pop(ignored, None)

        def compute(frame, **extra):
            # amep/evaluate.py:2219-2253
            c = frame.coords(ptype=ptype)
            hexorder = psi_k(c, frame.box, k=6)
            if other is None:
                return spatialcor(c, frame.box, hexorder, **kwargs, **extra)
            o = frame.coords(ptype=other)
            return spatialcor(c, frame.box, hexorder, other_coords=o, other_values=psi_k(o, frame.box, k=6),
                              **kwargs, **extra)

        res = self._run(traj, skip, nav, distributed, group, compute, allow_intra=True)
        self._avg, self._r = res[0], res[1]

    @property
    def r(self):
        return self._r
