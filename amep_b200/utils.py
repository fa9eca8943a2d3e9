"""Host helpers that mirror the pieces of amep/utils.py the hot path uses.

Only what RDF / S(q) need: frame selection (average_func), bin centres
(runningmean), unit vectors, the chunk-size rule of sf2d.  All of it is
parameter arithmetic on a handful of numbers; particle data never passes
through here.
"""
from __future__ import annotations

import os

import numpy as np

MAXMEM = 1.0  # GB per core: amep/base.py:176, enters the sf2d chunk rule


def available_cpu_count() -> int:
    """amep/utils.py:2253-2273."""
    if hasattr(os, "process_cpu_count"):
        return os.process_cpu_count()
    if hasattr(os, "sched_getaffinity"):
        return len(os.sched_getaffinity(0))
    return os.cpu_count()


def evaluated_indices(nframes: int, skip: float = 0.0, nr: int | None = 10) -> np.ndarray:
    """Frames picked by ``average_func`` (amep/utils.py:149-154):
    ``ceil(linspace(skip*N, N-1, nr))`` with ``nr`` clamped to the frames left."""
    n = nframes
    if nr is None or nr > n - skip * n:
        nr = max(1, int(n - skip * n))
    return np.array(np.ceil(np.linspace(skip * n, n - 1, nr)), dtype=int)


def runningmean(data: np.ndarray, nav: int, mode: str = "valid") -> np.ndarray:
    """amep/utils.py:190-233."""
    return np.convolve(data, np.ones((nav,)) / nav, mode=mode)


def optimal_chunksize(data_length: int, number_of_elements: int, bit_per_element: int = 32,
                      buffer: float = 0.25, maxmem: float = MAXMEM) -> int:
    """amep/utils.py:2024-2063 (only its value matters here: it fixes the
    particle chunking, hence the rounding sequence, of sf2d's complex64 sums)."""
    return int((maxmem - buffer) * 8e9 / bit_per_element / number_of_elements)


def unit_vector_2D(theta):
    """amep/utils.py:430-473 (angles in degrees)."""
    theta = theta * 2 * np.pi / 360
    if type(theta) == np.ndarray:
        return np.array([np.cos(theta), np.sin(theta), np.zeros(len(theta))]).T
    elif type(theta) == float:
        return np.array([np.cos(theta), np.sin(theta), 0])


def unit_vector_3D(theta, phi):
    """amep/utils.py:476-544 (angles in degrees)."""
    theta = theta * 2 * np.pi / 360
    phi = phi * 2 * np.pi / 360
    if type(theta) == np.ndarray and type(phi) == np.ndarray and len(theta) == len(phi):
        return np.array([np.sin(theta) * np.cos(phi), np.sin(theta) * np.sin(phi), np.cos(theta)]).T
    elif type(theta) == float and type(phi) == float:
        return np.array([np.sin(theta) * np.cos(phi), np.sin(theta) * np.sin(phi), np.cos(theta)])


def sq_from_sf2d(S, Qx, Qy):
    """Radial average of a 2D structure factor -> ``(S(q), q)`` (amep/utils.py:1255-1311):
    rings of width dq = grid spacing, mean of ``S`` over the grid points of each ring.
    A few hundred thousand grid points at most: host NumPy, like the other epilogues."""
    radius = np.sqrt(Qx ** 2 + Qy ** 2)
    dq = np.max([np.abs(Qx[0, 1] - Qx[0, 0]), np.abs(Qy[1, 0] - Qy[0, 0])])
    rings = np.arange(0, np.min([np.max(Qx), np.max(Qy)]), dq)
    weighted, edges = np.histogram(radius, weights=S, bins=rings, density=False)
    counts = np.histogram(radius, bins=rings)[0]
    return weighted / counts, runningmean(edges, 2)


def dimension(coords, verbose: bool = False) -> int:
    """``amep.utils.dimension`` (amep/utils.py:1617-1667) evaluated by the
    stats kernel (amepb_frame_dimension); shape errors are raised on the host."""
    from . import _core
    shape = tuple(coords.shape)
    if len(shape) != 2 or shape[1] not in (2, 3):
        raise ValueError(f"coordinates do not have a known shape. {shape}")
    if shape[1] == 2:
        return 2
    return int(_core.frame_dimensions(coords)[0])
