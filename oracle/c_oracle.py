"""ctypes front end of oracle/rdf_oracle.c (the C restatement of the reference path).

TEST INFRASTRUCTURE ONLY -- imported by tests/, ``__graft_entry__.smoke()`` and
the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``; never by
``amep_b200``.  Brute force O(M * N_img) like the reference, optionally with
OpenMP threads; used where the NumPy oracle would take too long.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

import amep_oracle as orc

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "librdf_oracle.so")
_lib = None


def build():
    subprocess.run(["make", "-C", _HERE], check=True, stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(_HERE, "rdf_oracle.c")
        if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
            build()
        L = ctypes.CDLL(_SO)
        vp, dp, lg = ctypes.c_void_p, ctypes.POINTER(ctypes.c_double), ctypes.c_long
        L.oracle_images.restype = lg
        L.oracle_images.argtypes = [vp, ctypes.c_int, lg, dp, ctypes.c_int, ctypes.c_int, vp]
        L.oracle_rdf_hist.restype = None
        L.oracle_rdf_hist.argtypes = [vp, ctypes.c_int, lg, vp, ctypes.c_int, lg, lg, dp, ctypes.c_int,
                                      vp, ctypes.c_int]
        L.oracle_sfiso_sums.restype = None
        L.oracle_sfiso_sums.argtypes = [vp, ctypes.c_int, lg, dp, dp, ctypes.c_int, dp, ctypes.c_int]
        L.oracle_sf2d_modes.restype = None
        L.oracle_sf2d_modes.argtypes = [vp, ctypes.c_int, lg, dp, dp, lg, lg, lg, vp, ctypes.c_int]
        L.oracle_max_threads.restype = ctypes.c_int
        _lib = L
    return _lib


def max_threads():
    return int(lib().oracle_max_threads())


def _prep(a):
    a = np.asarray(a)
    if a.dtype not in (np.float32, np.float64):
        a = a.astype(np.float64)
    return np.ascontiguousarray(a)


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def images(coords, box_boundary):
    """float32 periodic images, as oracle.periodic_images but in C."""
    c = _prep(coords)
    bb = np.asarray(box_boundary)
    dim = orc.dimension(c)
    if dim not in (2, 3):
        raise ValueError("amep.pbc.pbc_points: Only 2d or 3d systems are supported. "
                         "Please check coords and enforce_nd.")
    out = np.empty((27 * len(c), 3), dtype=np.float32)
    bb64 = np.ascontiguousarray(bb, dtype=np.float64)
    kept = lib().oracle_images(c.ctypes.data, int(c.dtype == np.float64), len(c), _dp(bb64),
                               int(bb.dtype == np.float32), dim, out.ctypes.data)
    return out[:kept]


def rdf_counts(coords, box_boundary, other_coords=None, nbins=None, rmax=None, pbc=True,
               nthreads=0, query_slice=None):
    """Integer pair histogram (same contract as amep_oracle.rdf_counts)."""
    c = _prep(coords)
    q = c if other_coords is None else _prep(other_coords)
    edges, nbins = orc.rdf_edges(np.asarray(box_boundary), nbins, rmax)
    e64 = np.ascontiguousarray(edges, dtype=np.float64)
    targets = images(c, box_boundary) if pbc else c
    hist = np.zeros(nbins, dtype=np.int64)
    qb, qe = (0, len(q)) if query_slice is None else query_slice
    lib().oracle_rdf_hist(targets.ctypes.data, int(targets.dtype == np.float64), len(targets),
                          q.ctypes.data, int(q.dtype == np.float64), qb, qe, _dp(e64), nbins,
                          hist.ctypes.data, nthreads)
    return hist, edges


def sfiso_fast(coords, box_boundary, qmax=20.0, twod=True, num=64, other_coords=None, nthreads=0):
    c = _prep(coords)
    o = c if other_coords is None else _prep(other_coords)
    q = np.ascontiguousarray(orc.sf_qvalues(box_boundary, qmax))
    dirs = orc.sfiso_directions(twod, num)
    total = 0
    for u in dirs:
        u = np.ascontiguousarray(u, dtype=np.float64)
        sa = np.empty((len(q), 2))
        lib().oracle_sfiso_sums(c.ctypes.data, int(c.dtype == np.float64), len(c), _dp(u), _dp(q), len(q),
                                _dp(sa), nthreads)
        if o is c:
            sb = sa
        else:
            sb = np.empty((len(q), 2))
            lib().oracle_sfiso_sums(o.ctypes.data, int(o.dtype == np.float64), len(o), _dp(u), _dp(q),
                                    len(q), _dp(sb), nthreads)
        total = total + (sa[:, 0] * sb[:, 0] + sa[:, 1] * sb[:, 1]) / np.sqrt(len(c) * len(o))
    return total / len(dirs), q


def sf2d_modes(coords, box_boundary, qmax, chunksize, nthreads=0, grid_slice=None):
    """complex64 rho(q) of sf2d(mode='std') (flattened grid slice -> complex64 array)."""
    c = _prep(coords)
    qx, qy = orc.sf2d_grid(box_boundary, qmax)
    fx = np.ascontiguousarray(qx.ravel())
    fy = np.ascontiguousarray(qy.ravel())
    gb, ge = (0, fx.size) if grid_slice is None else grid_slice
    out = np.empty((ge - gb, 2), dtype=np.float32)
    lib().oracle_sf2d_modes(c.ctypes.data, int(c.dtype == np.float64), len(c), _dp(fx), _dp(fy), gb, ge,
                            int(chunksize), out.ctypes.data, nthreads)
    rho = out[:, 0] + 1j * out[:, 1]
    return rho.astype(np.complex64), qx, qy


def sf2d_std(coords, box_boundary, other_coords=None, qmax=20.0, ntot=None, chunksize=None, nthreads=0):
    c = _prep(coords)
    same = other_coords is None
    o = c if same else _prep(other_coords)
    if ntot is None:
        ntot = len(c)
    cs = orc.sf2d_chunksize(len(c), len(o), chunksize)
    rho, qx, qy = sf2d_modes(c, box_boundary, qmax, cs, nthreads)
    rho = rho.reshape(qx.shape)
    rho_o = rho if same else sf2d_modes(o, box_boundary, qmax, cs, nthreads)[0].reshape(qx.shape)
    return np.real(rho * rho_o.conj()) / ntot, qx, qy
