"""The psi_6 orientation prologue of ``evaluate.SF2d(rotate=True)`` and ``evaluate.PCF2d``.

``psi_k`` mirrors ``amep.order.psi_k`` (amep/order.py:1361-1501) for its defaults
(``other_coords=None``, ``pbc=True``); the neighbour search, the bond sums and the rotated crop run
in libamepb's sm_100a kernels (csrc/order.cu).  What stays on the host are the few scalar
expressions of amep/evaluate.py:1352-1363 -- ``np.mean`` of the per-particle psi (NumPy's pairwise
complex sum, so that the mean has the reference's rounding), ``arccos`` of a normalised dot
product -- written as the reference writes them.
"""
from __future__ import annotations

import numpy as np

from . import _core

__all__ = ["psi_k", "psi6_mean", "psi6_orientation", "rotated_periodic_crop"]


def psi_k(coords, box_boundary, other_coords=None, rmax: float = 1.122, k: int = 6, pbc: bool = True):
    """k-atic bond order parameter of every particle of a 2D frame (amep/order.py:1361-1501).

    Only the reference's default neighbour set is on the B200 path (``other_coords=None``,
    ``pbc=True``); ``rmax`` is ignored exactly as in the reference."""
    if other_coords is not None or not pbc:
        raise NotImplementedError("amep_b200.order.psi_k: only other_coords=None, pbc=True (what evaluate.SF2d / "
                                  "PCF2d use, amep/evaluate.py:1352-1357, 730-735) is built")
    return _core.psi_k(coords, box_boundary, k=k)


def psi6_mean(coords, box_boundary):
    """``np.mean(psi_k(coords, box, k=6))`` -> ``(re, im)`` (amep/evaluate.py:730-735, 1352-1357)."""
    psi = np.mean(psi_k(coords, box_boundary, k=6))
    return np.array([psi.real, psi.imag])


def orientation_angle(psi):
    """Angle of the mean hexagonal order parameter ``psi = (re, im)`` (amep/evaluate.py:1359-1363)."""
    ex = np.array([1, 0])
    theta = np.arccos(np.dot(ex, psi) / np.sqrt(psi[0] ** 2 + psi[1] ** 2))
    if psi[1] < 0:
        theta = 2 * np.pi - theta
    return theta


def psi6_orientation(coords, box_boundary):
    """Mean sample orientation of one frame (amep/evaluate.py:1352-1363)."""
    return orientation_angle(psi6_mean(coords, box_boundary))


def rotated_periodic_crop(coords, box_boundary, theta):
    """``in_box(rotate_coords(pbc_points(coords, box), -theta, center), box)``
    (amep/evaluate.py:1366-1377): all periodic images rotated by ``-theta`` about the box centre,
    those inside the box kept, float64, in the reference's order.  NumPy in, NumPy out; CUDA
    tensor in, CUDA tensor out."""
    shape = tuple(coords.shape)
    if len(shape) != 2 or shape[1] != 3:
        raise ValueError(f"coordinates do not have a known shape. {shape}")
    dim = int(_core.frame_dimensions(coords)[0])   # pbc_points: enforce_nd=None -> dimension(coords)
    if dim not in (2, 3):
        raise ValueError("amep.pbc.pbc_points: Only 2d or 3d systems are supported. "
                         "Please check coords and enforce_nd.")
    # the entries of rotation(-theta) as NumPy computes them (amep/utils.py:575-577)
    return _core.rotate_crop(coords, box_boundary, dim, (np.cos(-theta), np.sin(-theta)))
