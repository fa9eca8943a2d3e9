"""Array plumbing between NumPy / PyTorch and the C ABI (no arithmetic here)."""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib


def _is_torch(x) -> bool:
    return type(x).__module__.split(".")[0] == "torch"


class Particles:
    """A batch of coordinate frames ``[F, N, 3]`` in host (NumPy) or device
    (CUDA tensor) memory, float32 or float64, C-contiguous."""

    def __init__(self, data, name="coords"):
        self.tensor = None
        self.array = None
        if _is_torch(data):
            import torch
            t = data
            if t.dtype not in (torch.float32, torch.float64):
                # NumPy promotion of the reference: integers behave like float64
                # in `c - coords` (amep/spatialcor.py:381), halves like float32
                t = t.to(torch.float32 if t.dtype in (torch.float16, torch.bfloat16) else torch.float64)
            if t.dim() == 2:
                t = t.unsqueeze(0)
            if t.dim() != 3 or t.shape[-1] != 3:
                raise ValueError(f"coordinates do not have a known shape. {tuple(data.shape)}")
            t = t.contiguous()
            if t.is_cuda:
                self.tensor = t
                self.dtype = _lib.F32 if t.dtype == torch.float32 else _lib.F64
                self.shape = tuple(t.shape)
                return
            data = t.numpy()
        a = np.asarray(data)
        if a.dtype not in (np.float32, np.float64):
            a = a.astype(np.float32 if a.dtype == np.float16 else np.float64)
        if a.ndim == 2:
            a = a[None]
        if a.ndim != 3 or a.shape[-1] != 3:
            raise ValueError(f"coordinates do not have a known shape. {np.shape(data)}")
        self.array = np.ascontiguousarray(a)
        self.dtype = _lib.F32 if self.array.dtype == np.float32 else _lib.F64
        self.shape = self.array.shape

    @property
    def on_device(self) -> bool:
        return self.tensor is not None

    @property
    def nframes(self) -> int:
        return self.shape[0]

    @property
    def n(self) -> int:
        return self.shape[1]

    def ptr(self) -> int:
        if self.tensor is not None:
            return self.tensor.data_ptr()
        return self.array.ctypes.data

    def to_device(self, device):
        import torch
        if self.tensor is not None:
            return self
        out = Particles.__new__(Particles)
        out.array = None
        out.tensor = torch.from_numpy(self.array).to(device, non_blocking=False)
        out.dtype, out.shape = self.dtype, self.shape
        return out


def box_array(box_boundary, nframes):
    """-> (float64 [F,3,2] contiguous, dtype code of the caller's box)."""
    if _is_torch(box_boundary):
        box_boundary = box_boundary.detach().cpu().numpy()
    b = np.asarray(box_boundary)
    code = _lib.F32 if b.dtype == np.float32 else _lib.F64
    if b.ndim == 2:
        b = np.broadcast_to(b, (nframes, 3, 2))
    if b.shape != (nframes, 3, 2):
        raise ValueError(f"box_boundary must have shape (3,2) or ({nframes},3,2), got {b.shape}")
    return np.ascontiguousarray(b, dtype=np.float64), code


def to_numpy(x):
    """Host NumPy view / copy of a NumPy array or (CUDA) tensor."""
    if _is_torch(x):
        return x.detach().cpu().numpy()
    return np.asarray(x)


def current_device() -> int:
    """The calling process' current CUDA device (one process per GPU sets it once with
    ``torch.cuda.set_device``); 0 when torch has no say."""
    try:
        import torch
        if torch.cuda.is_available():
            return int(torch.cuda.current_device())
    except Exception:  # pragma: no cover
        pass
    return 0


def _device_of(p: Particles, device):
    if p.on_device:
        return p.tensor.device.index if p.tensor.device.index is not None else 0
    if device is None:
        return current_device()
    if isinstance(device, int):
        return device
    import torch
    d = torch.device(device)
    return d.index if d.index is not None else 0


def _stream_for(dev_index, on_device):
    if not on_device:
        return None
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream(dev_index).cuda_stream)


def rdf_histograms(coords, box_boundary, edges, other=None, pbc=True, device=None, out=None, shard=None):
    """Integer pair histograms for a batch of frames (amepb_rdf_batch).

    Returns ``(hist, dims)``: ``hist`` int64 ``[F, nbins]`` (a CUDA tensor if the
    coordinates live on the device, else a NumPy array), ``dims`` int32 ``[F]``.
    Raises ValueError for frames that are neither 2D nor 3D (as the reference,
    amep/pbc.py:290-294).  ``shard=(index, count)``: evaluate only this share of every frame's
    query cell runs (amepb_rdf_set_shard); the histograms of all shares add up to the full one.
    """
    pc = coords if isinstance(coords, Particles) else Particles(coords)
    po = None
    if other is not None:
        po = other if isinstance(other, Particles) else Particles(other, "other_coords")
        if po.nframes != pc.nframes:
            raise ValueError("coords and other_coords must have the same number of frames")
        if pc.on_device != po.on_device:
            dev = (pc if pc.on_device else po).tensor.device
            pc, po = pc.to_device(dev), po.to_device(dev)
    nf = pc.nframes
    if pc.n == 0 or (po is not None and po.n == 0):
        raise IndexError("index 0 is out of bounds for axis 0 with size 0")
    box, box_code = box_array(box_boundary, nf)
    e = np.ascontiguousarray(np.asarray(edges), dtype=np.float64)
    if e.ndim == 1:
        nbins, stride = e.shape[0] - 1, 0
    elif e.ndim == 2 and e.shape[0] == nf:
        nbins, stride = e.shape[1] - 1, e.shape[1]
    else:
        raise ValueError("edges must have shape (nbins+1,) or (F, nbins+1)")
    if nbins < 1:
        raise ValueError("`bins` must be positive, when an integer")
    dev_index = _device_of(pc, device)
    ctx = _lib.context(dev_index)
    dims = np.zeros(nf, dtype=np.int32)
    if pc.on_device:
        import torch
        if out is None:
            out = torch.empty((nf, nbins), dtype=torch.int64, device=pc.tensor.device)
        hist_ptr = out.data_ptr()
        space = _lib.DEVICE
    else:
        if out is None:
            out = np.empty((nf, nbins), dtype=np.int64)
        hist_ptr = out.ctypes.data
        space = _lib.HOST
    if shard is not None:
        ctx.set_shard(*shard)
    try:
        rc = ctx.lib.amepb_rdf_batch(
            ctx.handle,
            ctypes.c_void_p(pc.ptr()), pc.dtype, pc.n,
            ctypes.c_void_p(po.ptr()) if po is not None else None,
            po.dtype if po is not None else 0, po.n if po is not None else 0,
            space, nf,
            box.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), box_code,
            e.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), stride, nbins,
            1 if pbc else 0,
            ctypes.c_void_p(hist_ptr), dims.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)),
            _stream_for(dev_index, pc.on_device))
    finally:
        if shard is not None:
            ctx.set_shard(0, 1)
    if rc == _lib.E_DIMENSION:
        raise ValueError(ctx.error())
    ctx.check(rc, "amepb_rdf_batch")
    return out, dims


def frame_dimensions(coords, device=None):
    """dimension() of every frame (amep/utils.py:1617-1667) -> int32 ``[F]``."""
    pc = coords if isinstance(coords, Particles) else Particles(coords)
    if pc.n == 0:
        raise IndexError("index 0 is out of bounds for axis 0 with size 0")
    dev_index = _device_of(pc, device)
    ctx = _lib.context(dev_index)
    dims = np.zeros(pc.nframes, dtype=np.int32)
    rc = ctx.lib.amepb_frame_dimension(
        ctx.handle, ctypes.c_void_p(pc.ptr()), pc.dtype,
        _lib.DEVICE if pc.on_device else _lib.HOST, pc.nframes, pc.n,
        dims.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), _stream_for(dev_index, pc.on_device))
    ctx.check(rc, "amepb_frame_dimension")
    return dims


def sq_harmonics(coords, kvec, nharm, device=None):
    """``out[f, d, h] = (sum cos((h+1) k_d.r), sum sin((h+1) k_d.r))`` (amepb_sq_harmonics).

    ``kvec``: float64 ``[ndir, 3]`` (shared) or ``[F, ndir, 3]``."""
    pc = coords if isinstance(coords, Particles) else Particles(coords)
    nf = pc.nframes
    k = np.ascontiguousarray(np.asarray(kvec), dtype=np.float64)
    if k.ndim == 2:
        ndir, stride = k.shape[0], 0
    elif k.ndim == 3 and k.shape[0] == nf:
        ndir, stride = k.shape[1], k.shape[1] * 3
    else:
        raise ValueError("kvec must have shape (ndir,3) or (F,ndir,3)")
    dev_index = _device_of(pc, device)
    ctx = _lib.context(dev_index)
    if pc.on_device:
        import torch
        out = torch.empty((nf, ndir, nharm, 2), dtype=torch.float64, device=pc.tensor.device)
        optr, space = out.data_ptr(), _lib.DEVICE
    else:
        out = np.empty((nf, ndir, nharm, 2), dtype=np.float64)
        optr, space = out.ctypes.data, _lib.HOST
    rc = ctx.lib.amepb_sq_harmonics(
        ctx.handle, ctypes.c_void_p(pc.ptr()), pc.dtype, space, nf, pc.n,
        k.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), stride, ndir, int(nharm),
        ctypes.c_void_p(optr), _stream_for(dev_index, pc.on_device))
    ctx.check(rc, "amepb_sq_harmonics")
    return out


def sq_debye(coords, other, q, twod, device=None):
    """Debye sums ``out[f, k] = sum_ij K(q_k d_ij)`` over all ordered pairs (amepb_sq_debye).

    ``q``: float64 ``[nq]`` (shared) or ``[F, nq]``; ``other`` may be None (= coords)."""
    pc = coords if isinstance(coords, Particles) else Particles(coords)
    po = None
    if other is not None:
        po = other if isinstance(other, Particles) else Particles(other)
        if pc.on_device != po.on_device:
            dev = (pc if pc.on_device else po).tensor.device
            pc, po = pc.to_device(dev), po.to_device(dev)
    nf = pc.nframes
    qa = np.ascontiguousarray(np.asarray(q, dtype=np.float64))
    if qa.ndim == 1:
        nq, stride = qa.shape[0], 0
    elif qa.ndim == 2 and qa.shape[0] == nf:
        nq, stride = qa.shape[1], qa.shape[1]
    else:
        raise ValueError("q must have shape (nq,) or (F, nq)")
    dev_index = _device_of(pc, device)
    ctx = _lib.context(dev_index)
    if pc.on_device:
        import torch
        out = torch.empty((nf, nq), dtype=torch.float64, device=pc.tensor.device)
        optr, space = out.data_ptr(), _lib.DEVICE
    else:
        out = np.empty((nf, nq), dtype=np.float64)
        optr, space = out.ctypes.data, _lib.HOST
    if nq == 0:
        return out
    rc = ctx.lib.amepb_sq_debye(
        ctx.handle, ctypes.c_void_p(pc.ptr()), pc.dtype, pc.n,
        ctypes.c_void_p(po.ptr()) if po is not None else None, po.dtype if po is not None else 0,
        po.n if po is not None else 0, space, nf,
        qa.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), stride, nq, 1 if twod else 0,
        ctypes.c_void_p(optr), _stream_for(dev_index, pc.on_device))
    ctx.check(rc, "amepb_sq_debye")
    return out


def sq_debye_ref(coords, other, q, twod, chunksize, device=None):
    """The Debye sums of ``sq_debye`` in the reference's float32 arithmetic and summation order
    (amepb_sq_debye_ref): float32 ``[F, nq]``, chunks of ``chunksize`` particles of ``other``."""
    pc = coords if isinstance(coords, Particles) else Particles(coords)
    po = None
    if other is not None:
        po = other if isinstance(other, Particles) else Particles(other)
        if pc.on_device != po.on_device:
            dev = (pc if pc.on_device else po).tensor.device
            pc, po = pc.to_device(dev), po.to_device(dev)
    nf = pc.nframes
    qa = np.ascontiguousarray(np.asarray(q, dtype=np.float64))
    if qa.ndim == 1:
        nq, stride = qa.shape[0], 0
    elif qa.ndim == 2 and qa.shape[0] == nf:
        nq, stride = qa.shape[1], qa.shape[1]
    else:
        raise ValueError("q must have shape (nq,) or (F, nq)")
    if int(chunksize) < 1:
        raise ValueError("chunksize must be positive")
    dev_index = _device_of(pc, device)
    ctx = _lib.context(dev_index)
    if pc.on_device:
        import torch
        out = torch.empty((nf, nq), dtype=torch.float32, device=pc.tensor.device)
        optr, space = out.data_ptr(), _lib.DEVICE
    else:
        out = np.empty((nf, nq), dtype=np.float32)
        optr, space = out.ctypes.data, _lib.HOST
    if nq == 0:
        return out
    rc = ctx.lib.amepb_sq_debye_ref(
        ctx.handle, ctypes.c_void_p(pc.ptr()), pc.dtype, pc.n,
        ctypes.c_void_p(po.ptr()) if po is not None else None, po.dtype if po is not None else 0,
        po.n if po is not None else 0, space, nf,
        qa.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), stride, nq, 1 if twod else 0, int(chunksize),
        ctypes.c_void_p(optr), _stream_for(dev_index, pc.on_device))
    ctx.check(rc, "amepb_sq_debye_ref")
    return out


def density_counts(coords, box_boundary, xedges, yedges, pbc=True, device=None):
    """Particle counts on the grid of ``amep.continuum.hde`` (amepb_density_counts): int32
    ``[F, ny, nx]`` (a CUDA tensor if the coordinates live on the device, else NumPy).

    ``xedges`` / ``yedges``: the reference's ``np.arange`` bin edges (any float dtype; widened to
    float64 exactly)."""
    pc = coords if isinstance(coords, Particles) else Particles(coords)
    nf = pc.nframes
    box, box_code = box_array(box_boundary, nf)
    xe = np.ascontiguousarray(np.asarray(xedges), dtype=np.float64)
    ye = np.ascontiguousarray(np.asarray(yedges), dtype=np.float64)
    if xe.ndim != 1 or ye.ndim != 1 or len(xe) < 2 or len(ye) < 2:
        raise ValueError("bin edges must be one-dimensional with at least two entries")
    dev_index = _device_of(pc, device)
    ctx = _lib.context(dev_index)
    shape = (nf, len(ye) - 1, len(xe) - 1)
    if pc.on_device:
        import torch
        out = torch.empty(shape, dtype=torch.int32, device=pc.tensor.device)
        optr, space = out.data_ptr(), _lib.DEVICE
    else:
        out = np.empty(shape, dtype=np.int32)
        optr, space = out.ctypes.data, _lib.HOST
    if nf == 0 or pc.n == 0:
        if pc.on_device:
            out.zero_()
        else:
            out[...] = 0
        return out
    rc = ctx.lib.amepb_density_counts(
        ctx.handle, ctypes.c_void_p(pc.ptr()), pc.dtype, pc.n, space, nf,
        box.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), box_code, 1 if pbc else 0,
        xe.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), len(xe),
        ye.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), len(ye),
        ctypes.c_void_p(optr), _stream_for(dev_index, pc.on_device))
    ctx.check(rc, "amepb_density_counts")
    return out


def pcf2d_counts(coords, other, center, center_f32, rot, xedges, yedges, device=None, query_offset=None):
    """Pair counts on the (dx, dy) grid of ``amep.spatialcor.pcf2d`` (amepb_pcf2d_counts): int64
    ``[F, nx, ny]`` (a CUDA tensor if the coordinates live on the device, else NumPy).

    ``other=None`` means the same particles (equal indices are skipped).  ``center``: float64
    ``[F, 3]``; ``rot``: ``None`` or float64 ``[F, 3]`` = (flag, cos, sin); ``xedges`` / ``yedges``:
    ``[nx+1]`` / ``[ny+1]`` shared by all frames, or ``[F, nx+1]`` / ``[F, ny+1]``."""
    pc = coords if isinstance(coords, Particles) else Particles(coords)
    same = other is None
    po = pc if same else (other if isinstance(other, Particles) else Particles(other, "other_coords"))
    nf = pc.nframes
    if po.nframes != nf:
        raise ValueError("coords and other_coords hold different numbers of frames")
    if po.on_device != pc.on_device:
        raise ValueError("coords and other_coords must both be NumPy arrays or both CUDA tensors")
    xe = np.asarray(xedges, dtype=np.float64)
    ye = np.asarray(yedges, dtype=np.float64)
    if xe.ndim != ye.ndim or xe.ndim not in (1, 2) or xe.shape[-1] < 2 or ye.shape[-1] < 2:
        raise ValueError("bin edges must be [nx+1] / [ny+1] or [F, nx+1] / [F, ny+1] with at least two entries")
    nxe, nye = xe.shape[-1], ye.shape[-1]
    stride = 0
    if xe.ndim == 2:
        if len(xe) != nf or len(ye) != nf:
            raise ValueError("per-frame bin edges need one row per frame")
        stride = max(nxe, nye)
        xpad = np.zeros((nf, stride)); xpad[:, :nxe] = xe
        ypad = np.zeros((nf, stride)); ypad[:, :nye] = ye
        xe, ye = xpad, ypad
    xe = np.ascontiguousarray(xe)
    ye = np.ascontiguousarray(ye)
    cen = np.ascontiguousarray(np.broadcast_to(np.asarray(center, dtype=np.float64), (nf, 3)))
    rot_ptr = None
    if rot is not None:
        rot = np.ascontiguousarray(np.broadcast_to(np.asarray(rot, dtype=np.float64), (nf, 3)))
        rot_ptr = rot.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
    dev_index = _device_of(pc, device)
    ctx = _lib.context(dev_index)
    shape = (nf, nxe - 1, nye - 1)
    if pc.on_device:
        import torch
        out = torch.empty(shape, dtype=torch.int64, device=pc.tensor.device)
        optr, space = out.data_ptr(), _lib.DEVICE
    else:
        out = np.empty(shape, dtype=np.int64)
        optr, space = out.ctypes.data, _lib.HOST
    if nf == 0 or pc.n == 0 or po.n == 0:
        if pc.on_device:
            out.zero_()
        else:
            out[...] = 0
        return out
    # query_offset: ``other`` is the slice of ``coords`` from that particle on (intra-frame sharding,
    # amepb_pcf2d_set_query_offset): equal global indices are skipped like for other=None
    if query_offset is not None:
        ctx.set_pcf2d_query_offset(query_offset)
    try:
        rc = ctx.lib.amepb_pcf2d_counts(
            ctx.handle, ctypes.c_void_p(pc.ptr()), pc.dtype, pc.n, ctypes.c_void_p(po.ptr()), po.dtype, po.n,
            1 if same else 0, space, nf, cen.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), 1 if center_f32 else 0,
            rot_ptr, xe.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), nxe,
            ye.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), nye, stride,
            ctypes.c_void_p(optr), _stream_for(dev_index, pc.on_device))
        ctx.check(rc, "amepb_pcf2d_counts")
    finally:
        if query_offset is not None:
            ctx.set_pcf2d_query_offset(-1)
    return out


def sf2d_modes(coords, qpos, nq, accum="c64", chunksize=0, device=None, shard=None):
    """rho(q) on the sf2d grid (amepb_sf2d_modes) -> complex array ``[F, 2nq, 2nq]``.

    ``qpos``: the positive half of the q axis, float64 ``[nq]`` (shared) or ``[F, nq]``.
    ``shard=(index, count)``: only this share of the rows of q tiles, zeros elsewhere
    (amepb_sf2d_set_shard); the arrays of all shares add up to the full one exactly."""
    pc = coords if isinstance(coords, Particles) else Particles(coords)
    nf = pc.nframes
    q = np.ascontiguousarray(np.asarray(qpos, dtype=np.float64))
    if q.ndim == 1 and q.shape[0] == nq:
        stride = 0
    elif q.ndim == 2 and q.shape == (nf, nq):
        stride = nq
    else:
        raise ValueError("qpos must have shape (nq,) or (F, nq)")
    mode = {"c64": _lib.ACCUM_C64_REFERENCE, "f64": _lib.ACCUM_F64}[accum]
    dev_index = _device_of(pc, device)
    ctx = _lib.context(dev_index)
    g = 2 * int(nq)
    if pc.on_device:
        import torch
        out = torch.empty((nf, g, g), dtype=torch.complex64 if accum == "c64" else torch.complex128,
                          device=pc.tensor.device)
        optr, space = out.data_ptr(), _lib.DEVICE
    else:
        out = np.empty((nf, g, g), dtype=np.complex64 if accum == "c64" else np.complex128)
        optr, space = out.ctypes.data, _lib.HOST
    if nq == 0:
        return out
    if shard is not None:
        ctx.set_sf2d_shard(*shard)
    try:
        rc = ctx.lib.amepb_sf2d_modes(
            ctx.handle, ctypes.c_void_p(pc.ptr()), pc.dtype, space, nf, pc.n,
            q.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), stride, int(nq), mode, int(chunksize),
            ctypes.c_void_p(optr), _stream_for(dev_index, pc.on_device))
        ctx.check(rc, "amepb_sf2d_modes")
    finally:
        if shard is not None:
            ctx.set_sf2d_shard(0, 1)
    return out


def psi_k(coords, box_boundary, k=6, want_neighbors=False, device=None):
    """``amep.order.psi_k`` of one 2D frame (amepb_psi_k) -> complex array ``[n]`` (complex64 for
    float32 coordinates, like NumPy's arithmetic in the reference), optionally the neighbour ids
    ``[n, k]``.  Raises ValueError where the reference does: a particle with an image of itself
    among its neighbours makes the reference's id array ragged (amep/order.py:1340-1348, 1486)."""
    pc = coords if isinstance(coords, Particles) else Particles(coords)
    if pc.nframes != 1:
        raise ValueError("psi_k takes one frame")
    n = pc.n
    box, box_code = box_array(box_boundary, 1)
    dev_index = _device_of(pc, device)
    ctx = _lib.context(dev_index)
    f32 = pc.dtype == _lib.F32
    if pc.on_device:
        import torch
        out = torch.empty((n, 2), dtype=torch.float32 if f32 else torch.float64, device=pc.tensor.device)
        nb = torch.empty((n, k), dtype=torch.int32, device=pc.tensor.device) if want_neighbors else None
        optr, nptr, space = out.data_ptr(), (nb.data_ptr() if nb is not None else None), _lib.DEVICE
    else:
        out = np.empty((n, 2), dtype=np.float32 if f32 else np.float64)
        nb = np.empty((n, k), dtype=np.int32) if want_neighbors else None
        optr, nptr, space = out.ctypes.data, (nb.ctypes.data if nb is not None else None), _lib.HOST
    bad = ctypes.c_int64(0)
    rc = ctx.lib.amepb_psi_k(ctx.handle, ctypes.c_void_p(pc.ptr()), pc.dtype, n, space,
                             box.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), box_code, int(k),
                             ctypes.c_void_p(optr), ctypes.c_void_p(nptr) if nptr else None, ctypes.byref(bad),
                             _stream_for(dev_index, pc.on_device))
    ctx.check(rc, "amepb_psi_k")
    if bad.value:
        raise ValueError("setting an array element with a sequence. The requested array has an inhomogeneous shape: "
                         f"{bad.value} particle(s) have a periodic image of themselves (or fewer than k images) among "
                         "their k nearest neighbours (amep/order.py:1340-1348)")
    host = to_numpy(out)
    psi = host.view(np.complex64 if f32 else np.complex128).reshape(n)
    if want_neighbors:
        return psi, to_numpy(nb).astype(np.int64)
    return psi


def rotate_crop(coords, box_boundary, dim, cos_sin, device=None):
    """``in_box(rotate_coords(pbc_points(coords, box), angle, center), box)`` of one frame
    (amepb_rotate_crop) -> float64 ``[m, 3]`` in the memory space of ``coords``."""
    pc = coords if isinstance(coords, Particles) else Particles(coords)
    if pc.nframes != 1:
        raise ValueError("rotate_crop takes one frame")
    n = pc.n
    box, box_code = box_array(box_boundary, 1)
    dev_index = _device_of(pc, device)
    ctx = _lib.context(dev_index)
    rot = np.ascontiguousarray(np.asarray(cos_sin, dtype=np.float64).reshape(2))
    count = ctypes.c_int64(0)
    space = _lib.DEVICE if pc.on_device else _lib.HOST
    stream = _stream_for(dev_index, pc.on_device)
    args = (ctx.handle, ctypes.c_void_p(pc.ptr()), pc.dtype, n, space,
            box.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), box_code, int(dim),
            rot.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
    rc = ctx.lib.amepb_rotate_crop(*args, None, 0, ctypes.byref(count), stream)
    ctx.check(rc, "amepb_rotate_crop")
    m = int(count.value)
    if pc.on_device:
        import torch
        out = torch.empty((m, 3), dtype=torch.float64, device=pc.tensor.device)
        optr = out.data_ptr()
    else:
        out = np.empty((m, 3), dtype=np.float64)
        optr = out.ctypes.data
    if m:
        rc = ctx.lib.amepb_rotate_crop(*args, ctypes.c_void_p(optr), m, ctypes.byref(count), stream)
        ctx.check(rc, "amepb_rotate_crop")
    return out


def _single(p: Particles, what: str):
    if p.nframes != 1:
        raise ValueError(f"{what} takes one frame")
    return p


def pcf_angle_counts(coords, other, same, box_boundary, pbc, e, angle, dedges, aedges, device=None, query_offset=None):
    """Pair counts on the (distance, angle) grid of ``amep.spatialcor.pcf_angle``
    (amepb_pcf_angle_counts) -> int64 ``[nd, na]`` NumPy array."""
    pc = _single(coords if isinstance(coords, Particles) else Particles(coords), "pcf_angle")
    po = pc if other is None else _single(other if isinstance(other, Particles) else Particles(other, "other_coords"), "pcf_angle")
    if pc.on_device != po.on_device:
        dev = (pc if pc.on_device else po).tensor.device
        pc, po = pc.to_device(dev), po.to_device(dev)
    box, box_code = box_array(box_boundary, 1)
    de = np.ascontiguousarray(np.asarray(dedges), dtype=np.float64)
    ae = np.ascontiguousarray(np.asarray(aedges), dtype=np.float64)
    e = np.asarray(e, dtype=np.float64)
    if e.ndim == 1:
        e_arr, stride = np.ascontiguousarray(e), 0
        e_norm = np.array([np.linalg.norm(e)], dtype=np.float64)
    else:
        e_arr, stride = np.ascontiguousarray(e), 3
        if e_arr.shape != (po.n, 3):
            raise ValueError("e must have shape (3,) or one direction per particle of other_coords")
        e_norm = np.ascontiguousarray(np.sqrt((e_arr[:, 0] * e_arr[:, 0] + e_arr[:, 1] * e_arr[:, 1]) + e_arr[:, 2] * e_arr[:, 2]))
    dev_index = _device_of(pc, device)
    ctx = _lib.context(dev_index)
    shape = (len(de) - 1, len(ae) - 1)
    if pc.on_device:
        import torch
        out = torch.empty(shape, dtype=torch.int64, device=pc.tensor.device)
        optr, space = out.data_ptr(), _lib.DEVICE
    else:
        out = np.empty(shape, dtype=np.int64)
        optr, space = out.ctypes.data, _lib.HOST
    dp = ctypes.POINTER(ctypes.c_double)
    # query_offset: ``other`` (and ``e``) is the slice of ``coords`` from that particle on (intra-frame sharding)
    if query_offset is not None:
        ctx.set_pcf_angle_query_offset(query_offset)
    try:
        rc = ctx.lib.amepb_pcf_angle_counts(
            ctx.handle, ctypes.c_void_p(pc.ptr()), pc.dtype, pc.n, ctypes.c_void_p(po.ptr()), po.dtype, po.n,
            1 if same else 0, space, box.ctypes.data_as(dp), box_code, 1 if pbc else 0,
            e_arr.ctypes.data_as(dp), stride, e_norm.ctypes.data_as(dp), float(angle),
            de.ctypes.data_as(dp), len(de), ae.ctypes.data_as(dp), len(ae), ctypes.c_void_p(optr),
            _stream_for(dev_index, pc.on_device))
        ctx.check(rc, "amepb_pcf_angle_counts")
    finally:
        if query_offset is not None:
            ctx.set_pcf_angle_query_offset(-1)
    return to_numpy(out)


_VALUE_KINDS = {np.dtype(np.float32): 0, np.dtype(np.float64): 1, np.dtype(np.complex64): 2, np.dtype(np.complex128): 3}


def _values_array(values, n, on_device, device):
    """Per-particle values ``[n]`` or ``[n, d]`` -> (contiguous array / tensor in the space of the
    coordinates, kind code, components)."""
    if _is_torch(values):
        import torch
        v = values
        if v.dtype not in (torch.float32, torch.float64, torch.complex64, torch.complex128):
            v = v.to(torch.float64)
        v = v.contiguous()
        if on_device and not v.is_cuda:
            v = v.to(device)
        if not on_device:
            return _values_array(v.cpu().numpy(), n, False, device)
        kind = {torch.float32: 0, torch.float64: 1, torch.complex64: 2, torch.complex128: 3}[v.dtype]
        shape = tuple(v.shape)
        holder, ptr = v, v.data_ptr()
    else:
        a = np.asarray(values)
        if a.dtype not in _VALUE_KINDS:
            a = a.astype(np.complex128 if np.iscomplexobj(a) else np.float64)
        a = np.ascontiguousarray(a)
        if on_device:
            import torch
            return _values_array(torch.from_numpy(a).to(device), n, True, device)
        kind, shape, holder, ptr = _VALUE_KINDS[a.dtype], a.shape, a, a.ctypes.data
    if len(shape) not in (1, 2) or shape[0] != n:
        raise ValueError(f"amep.spatialcor.spatialcor: values have an invalid shape {shape}. The shape must be "
                         "(N,) for scalar values and (N,d) for vector quantities.")
    return holder, ptr, kind, (1 if len(shape) == 1 else shape[1])


def spatialcor_sums(coords, values, other, other_values, box_boundary, pbc, edges, device=None):
    """Pair counts and sums of ``conj(other_value) * value`` per distance bin of
    ``amep.spatialcor.spatialcor`` (amepb_spatialcor_sums) -> (norm int64 ``[nb]``, cor complex128 ``[nb]``)."""
    pc = _single(coords if isinstance(coords, Particles) else Particles(coords), "spatialcor")
    po = None if other is None else _single(other if isinstance(other, Particles) else Particles(other, "other_coords"), "spatialcor")
    if po is not None and pc.on_device != po.on_device:
        dev = (pc if pc.on_device else po).tensor.device
        pc, po = pc.to_device(dev), po.to_device(dev)
    tdev = pc.tensor.device if pc.on_device else None
    vh, vptr, kind, ncomp = _values_array(values, pc.n, pc.on_device, tdev)
    wh = wptr = None
    if po is not None:
        wh, wptr, kind_o, ncomp_o = _values_array(other_values, po.n, pc.on_device, tdev)
        if kind_o != kind or ncomp_o != ncomp:
            # NumPy would promote the product: bring both to the wider kind
            target = [np.float32, np.float64, np.complex64, np.complex128][max(kind, kind_o) if (kind < 2) == (kind_o < 2) else 3]
            vh, vptr, kind, ncomp = _values_array(to_numpy(vh).astype(target), pc.n, pc.on_device, tdev)
            wh, wptr, kind_o, ncomp_o = _values_array(to_numpy(wh).astype(target), po.n, pc.on_device, tdev)
            if ncomp_o != ncomp:
                raise ValueError("values and other_values must have the same number of components")
    if ncomp > 4:
        raise ValueError("amep_b200.spatialcor.spatialcor: at most 4 value components per particle")
    box, box_code = box_array(box_boundary, 1)
    e = np.ascontiguousarray(np.asarray(edges), dtype=np.float64)
    nb = len(e) - 1
    dev_index = _device_of(pc, device)
    ctx = _lib.context(dev_index)
    if pc.on_device:
        import torch
        norm = torch.empty(nb, dtype=torch.int64, device=pc.tensor.device)
        cor = torch.empty((nb, 2), dtype=torch.float64, device=pc.tensor.device)
        nptr, cptr, space = norm.data_ptr(), cor.data_ptr(), _lib.DEVICE
    else:
        norm = np.empty(nb, dtype=np.int64)
        cor = np.empty((nb, 2), dtype=np.float64)
        nptr, cptr, space = norm.ctypes.data, cor.ctypes.data, _lib.HOST
    dp = ctypes.POINTER(ctypes.c_double)
    rc = ctx.lib.amepb_spatialcor_sums(
        ctx.handle, ctypes.c_void_p(pc.ptr()), pc.dtype, pc.n,
        ctypes.c_void_p(po.ptr()) if po is not None else None, po.dtype if po is not None else 0,
        po.n if po is not None else 0, space, box.ctypes.data_as(dp), box_code, 1 if pbc else 0,
        ctypes.c_void_p(vptr), ctypes.c_void_p(wptr) if wptr else None, kind, ncomp,
        e.ctypes.data_as(dp), len(e), ctypes.c_void_p(nptr), ctypes.c_void_p(cptr),
        _stream_for(dev_index, pc.on_device))
    ctx.check(rc, "amepb_spatialcor_sums")
    c = to_numpy(cor)
    return to_numpy(norm), c[:, 0] + 1j * c[:, 1], kind >= 2
