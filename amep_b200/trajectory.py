"""Tensor-backed stand-in for ``amep.trajectory.ParticleTrajectory``.

The reference's trajectory is an HDF5 reader (amep/base.py:1263, amep/trajectory.py:52);
its storage engine is out of scope.  What the evaluate
classes need of it is small -- ``len(traj)``, ``traj[i]`` / ``traj[index array]``,
``traj.times`` and, per frame, ``coords(ptype=)``, ``box``, ``n()``, ``center``
(amep/base.py:429-1080) -- and this class offers exactly that over arrays that
already sit in memory: NumPy arrays or torch tensors, optionally resident in HBM
so that the kernels read them in place.
"""
from __future__ import annotations

import numpy as np

from . import _core


class TensorFrame:
    """One frame of a TensorTrajectory (duck-types amep.base.BaseFrame)."""

    def __init__(self, traj: "TensorTrajectory", index: int):
        self._traj = traj
        self._index = int(index)

    @property
    def step(self):
        return self._traj.steps[self._index]

    @property
    def time(self):
        return self._traj.times[self._index]

    @property
    def box(self):
        b = self._traj._box
        return b[self._index] if b.ndim == 3 else b

    @property
    def center(self):
        return np.mean(self.box, axis=1)

    def _mask(self, ptype):
        types = self._traj._types
        if ptype is None or types is None:
            return None
        t = types[self._index] if types.ndim == 2 else types
        wanted = np.atleast_1d(np.asarray(ptype))
        return np.isin(t, wanted)

    def coords(self, ptype=None, **kwargs):
        c = self._traj._coords[self._index]
        mask = self._mask(ptype)
        if mask is None:
            return c
        if _core._is_torch(c):
            import torch
            return c[torch.from_numpy(mask).to(c.device)]
        return c[mask]

    def _per_particle(self, data, what, ptype):
        if data is None:
            raise KeyError(f"The trajectory holds no {what}.")
        d = data[self._index]
        mask = self._mask(ptype)
        if mask is None:
            return d
        if _core._is_torch(d):
            import torch
            return d[torch.from_numpy(mask).to(d.device)]
        return d[mask]

    def velocities(self, ptype=None, **kwargs):
        return self._per_particle(self._traj._velocities, "velocities", ptype)

    def orientations(self, ptype=None, **kwargs):
        return self._per_particle(self._traj._orientations, "orientations", ptype)

    def n(self, ptype=None):
        mask = self._mask(ptype)
        return int(self._traj._coords.shape[1] if mask is None else mask.sum())


class TensorTrajectory:
    """Trajectory over in-memory arrays.

    coords : ``[F, N, 3]`` NumPy array or torch tensor (CPU or CUDA), float32 or float64
    box    : ``(3,2)`` or ``(F,3,2)`` array (its dtype is kept: it decides the
             reference's arithmetic)
    times, steps : optional ``[F]``;  types : optional ``[N]`` or ``[F, N]`` particle types
    velocities, orientations : optional ``[F, N, 3]`` (``frame.velocities()`` / ``frame.orientations()``,
             read by SpatialVelCor and PCFangle(mode='orientations'))
    """

    def __init__(self, coords, box, times=None, steps=None, types=None, velocities=None, orientations=None):
        if len(coords.shape) != 3 or coords.shape[-1] != 3:
            raise ValueError(f"coords must have shape (F, N, 3), got {tuple(coords.shape)}")
        self._coords = coords
        if _core._is_torch(box):
            box = box.detach().cpu().numpy()
        self._box = np.asarray(box)
        nf = coords.shape[0]
        self.steps = np.arange(nf) if steps is None else np.asarray(steps)
        self.times = self.steps.astype(float) if times is None else np.asarray(times)
        self._types = None if types is None else np.asarray(types)
        for name, data in (("velocities", velocities), ("orientations", orientations)):
            if data is not None and tuple(data.shape) != tuple(coords.shape):
                raise ValueError(f"{name} must have the shape of coords, {tuple(coords.shape)}")
        self._velocities = velocities
        self._orientations = orientations

    def __len__(self):
        return int(self._coords.shape[0])

    def __getitem__(self, item):
        if isinstance(item, slice):
            return [TensorFrame(self, i) for i in range(*item.indices(len(self)))]
        if isinstance(item, (list, tuple, np.ndarray)):
            return [TensorFrame(self, i) for i in item]
        i = int(item)
        if i < 0:
            i += len(self)
        if not 0 <= i < len(self):
            raise IndexError("frame index out of range")
        return TensorFrame(self, i)

    def __iter__(self):
        for i in range(len(self)):
            yield self[i]

    @property
    def on_device(self):
        return _core._is_torch(self._coords) and self._coords.is_cuda

    # fast path used by the evaluate classes: one gather instead of F frame objects
    def stack(self, indices, ptype=None):
        """-> (coords ``[len(indices), n, 3]``, box ``(3,2)`` or ``[len(indices),3,2]``)."""
        idx = np.asarray(indices, dtype=np.int64)
        c = self._coords
        contiguous = len(idx) > 0 and np.array_equal(idx, np.arange(idx[0], idx[0] + len(idx)))
        if _core._is_torch(c):
            import torch
            sel = c[int(idx[0]):int(idx[0]) + len(idx)] if contiguous else \
                c.index_select(0, torch.from_numpy(idx).to(c.device))
        else:
            sel = c[int(idx[0]):int(idx[0]) + len(idx)] if contiguous else c[idx]
        if ptype is not None and self._types is not None:
            if self._types.ndim == 2:
                raise ValueError("per-frame particle types: gather frame by frame")
            mask = np.isin(self._types, np.atleast_1d(np.asarray(ptype)))
            if _core._is_torch(sel):
                import torch
                sel = sel[:, torch.from_numpy(mask).to(sel.device)]
            else:
                sel = sel[:, mask]
        box = self._box[idx] if self._box.ndim == 3 else self._box
        return sel, box
