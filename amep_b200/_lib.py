"""ctypes binding of libamepb.so (the C ABI declared in include/amepb.h).

There is no CPU fallback: if the library is not built, or no B200-class CUDA
device is present, every compute entry point raises ``AmepbError``.
"""
from __future__ import annotations

import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("AMEPB_LIB", os.path.join(_HERE, "lib", "libamepb.so"))

F32, F64 = 0, 1
DEVICE, HOST = 0, 1
E_INVAL, E_DIMENSION, E_NOMEM, E_NODEVICE = -1, -2, -3, -4
ACCUM_C64_REFERENCE, ACCUM_F64 = 0, 1
ABI_VERSION = 1

# every symbol include/amepb.h declares
SYMBOLS = (
    "amepb_abi_version", "amepb_create", "amepb_destroy", "amepb_last_error",
    "amepb_frame_dimension", "amepb_rdf_batch", "amepb_rdf_last_info",
    "amepb_rdf_set_cells_per_cutoff", "amepb_rdf_set_symmetric", "amepb_rdf_set_shard", "amepb_sq_harmonics", "amepb_sq_debye", "amepb_sq_debye_ref", "amepb_sf2d_modes", "amepb_sf2d_set_shard",
    "amepb_density_counts",
    "amepb_pcf2d_counts", "amepb_pcf2d_set_query_offset", "amepb_psi_k", "amepb_rotate_crop", "amepb_pcf_angle_counts", "amepb_pcf_angle_set_query_offset", "amepb_spatialcor_sums",
    "amepb_kernel_launches", "amepb_synchronize", "amepb_set_timing", "amepb_last_kernel_ms",
)


class AmepbError(RuntimeError):
    """libamepb is missing, has no device, or a call failed."""


class RdfInfo(ctypes.Structure):
    _fields_ = [
        ("frames", ctypes.c_int64),
        ("images", ctypes.c_int64),
        ("cells", ctypes.c_int64),
        ("block_tasks", ctypes.c_int64),
        ("candidate_pairs", ctypes.c_int64),
        ("kernel_launches", ctypes.c_int64),
        ("cells_per_cutoff", ctypes.c_int32),
        ("sub_batches", ctypes.c_int32),
        ("pair_kernel_ms", ctypes.c_double),
        ("build_ms", ctypes.c_double),
    ]

    def as_dict(self):
        return {name: (float(getattr(self, name)) if typ is ctypes.c_double else int(getattr(self, name)))
                for name, typ in self._fields_}


_lib = None
_lib_lock = threading.Lock()


def load():
    """Load libamepb.so and declare its prototypes.  Raises AmepbError if absent."""
    global _lib
    if _lib is not None:
        return _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise AmepbError(
                f"{LIB_PATH} not found: build it with `make -C amep_b200/csrc` "
                "(or `python -c 'import __graft_entry__ as g; g.build()'`). "
                "amep_b200 has no CPU fallback.")
        try:
            lib = ctypes.CDLL(LIB_PATH)
        except OSError as exc:  # pragma: no cover
            raise AmepbError(f"cannot load {LIB_PATH}: {exc}") from exc
        c = ctypes
        vp, i32, i64 = c.c_void_p, c.c_int32, c.c_int64
        lib.amepb_abi_version.restype = c.c_int
        lib.amepb_abi_version.argtypes = []
        lib.amepb_create.restype = c.c_int
        lib.amepb_create.argtypes = [c.c_int, c.POINTER(vp)]
        lib.amepb_destroy.restype = None
        lib.amepb_destroy.argtypes = [vp]
        lib.amepb_last_error.restype = c.c_char_p
        lib.amepb_last_error.argtypes = [vp]
        lib.amepb_frame_dimension.restype = c.c_int
        lib.amepb_frame_dimension.argtypes = [vp, vp, c.c_int, c.c_int, i64, i64, c.POINTER(i32), vp]
        lib.amepb_rdf_batch.restype = c.c_int
        lib.amepb_rdf_batch.argtypes = [
            vp,                     # ctx
            vp, c.c_int, i64,       # coords, dtype, n
            vp, c.c_int, i64,       # other, dtype, m
            c.c_int, i64,           # space, nframes
            c.POINTER(c.c_double), c.c_int,        # box, box dtype
            c.POINTER(c.c_double), i64, i32,       # edges, stride, nbins
            c.c_int,                # pbc
            vp, c.POINTER(i32), vp  # hist_out, dim_out, stream
        ]
        lib.amepb_rdf_last_info.restype = c.c_int
        lib.amepb_rdf_last_info.argtypes = [vp, c.POINTER(RdfInfo)]
        lib.amepb_rdf_set_cells_per_cutoff.restype = c.c_int
        lib.amepb_rdf_set_cells_per_cutoff.argtypes = [vp, c.c_int]
        lib.amepb_rdf_set_symmetric.restype = c.c_int
        lib.amepb_rdf_set_symmetric.argtypes = [vp, c.c_int]
        lib.amepb_rdf_set_shard.restype = c.c_int
        lib.amepb_rdf_set_shard.argtypes = [vp, c.c_int, c.c_int]
        lib.amepb_pcf_angle_set_query_offset.restype = c.c_int
        lib.amepb_pcf_angle_set_query_offset.argtypes = [vp, i64]
        lib.amepb_pcf2d_set_query_offset.restype = c.c_int
        lib.amepb_pcf2d_set_query_offset.argtypes = [vp, i64]
        lib.amepb_sf2d_set_shard.restype = c.c_int
        lib.amepb_sf2d_set_shard.argtypes = [vp, c.c_int, c.c_int]
        lib.amepb_sq_harmonics.restype = c.c_int
        lib.amepb_sq_harmonics.argtypes = [vp, vp, c.c_int, c.c_int, i64, i64,
                                           c.POINTER(c.c_double), i64, i32, i32, vp, vp]
        lib.amepb_sq_debye.restype = c.c_int
        lib.amepb_sq_debye.argtypes = [vp, vp, c.c_int, i64, vp, c.c_int, i64, c.c_int, i64,
                                       c.POINTER(c.c_double), i64, i32, c.c_int, vp, vp]
        lib.amepb_sq_debye_ref.restype = c.c_int
        lib.amepb_sq_debye_ref.argtypes = [vp, vp, c.c_int, i64, vp, c.c_int, i64, c.c_int, i64,
                                           c.POINTER(c.c_double), i64, i32, c.c_int, i64, vp, vp]
        lib.amepb_sf2d_modes.restype = c.c_int
        lib.amepb_sf2d_modes.argtypes = [vp, vp, c.c_int, c.c_int, i64, i64,
                                         c.POINTER(c.c_double), i64, i32, c.c_int, i64, vp, vp]
        lib.amepb_density_counts.restype = c.c_int
        lib.amepb_density_counts.argtypes = [vp, vp, c.c_int, i64, c.c_int, i64, c.POINTER(c.c_double), c.c_int, c.c_int,
                                             c.POINTER(c.c_double), i32, c.POINTER(c.c_double), i32, vp, vp]
        lib.amepb_pcf2d_counts.restype = c.c_int
        lib.amepb_pcf2d_counts.argtypes = [vp, vp, c.c_int, i64, vp, c.c_int, i64, c.c_int, c.c_int, i64,
                                           c.POINTER(c.c_double), c.c_int, c.POINTER(c.c_double),
                                           c.POINTER(c.c_double), i32, c.POINTER(c.c_double), i32, i64, vp, vp]
        dp = c.POINTER(c.c_double)
        lib.amepb_pcf_angle_counts.restype = c.c_int
        lib.amepb_pcf_angle_counts.argtypes = [vp, vp, c.c_int, i64, vp, c.c_int, i64, c.c_int, c.c_int, dp, c.c_int, c.c_int,
                                               dp, i64, dp, c.c_double, dp, i32, dp, i32, vp, vp]
        lib.amepb_spatialcor_sums.restype = c.c_int
        lib.amepb_spatialcor_sums.argtypes = [vp, vp, c.c_int, i64, vp, c.c_int, i64, c.c_int, dp, c.c_int, c.c_int,
                                              vp, vp, c.c_int, i32, dp, i32, vp, vp, vp]
        lib.amepb_psi_k.restype = c.c_int
        lib.amepb_psi_k.argtypes = [vp, vp, c.c_int, i64, c.c_int, c.POINTER(c.c_double), c.c_int, i32, vp, vp,
                                    c.POINTER(i64), vp]
        lib.amepb_rotate_crop.restype = c.c_int
        lib.amepb_rotate_crop.argtypes = [vp, vp, c.c_int, i64, c.c_int, c.POINTER(c.c_double), c.c_int, c.c_int,
                                          c.POINTER(c.c_double), vp, i64, c.POINTER(i64), vp]
        lib.amepb_kernel_launches.restype = i64
        lib.amepb_kernel_launches.argtypes = [vp]
        lib.amepb_synchronize.restype = c.c_int
        lib.amepb_synchronize.argtypes = [vp, vp]
        lib.amepb_set_timing.restype = c.c_int
        lib.amepb_set_timing.argtypes = [vp, c.c_int]
        lib.amepb_last_kernel_ms.restype = c.c_double
        lib.amepb_last_kernel_ms.argtypes = [vp]
        if lib.amepb_abi_version() != ABI_VERSION:
            raise AmepbError(f"{LIB_PATH}: ABI version {lib.amepb_abi_version()} != {ABI_VERSION}")
        _lib = lib
    return _lib


class Context:
    """One libamepb context (device workspaces).  Not shared between threads."""

    def __init__(self, device: int = 0):
        self.lib = load()
        handle = ctypes.c_void_p()
        rc = self.lib.amepb_create(int(device), ctypes.byref(handle))
        if rc != 0:
            msg = self.lib.amepb_last_error(None).decode()
            raise AmepbError(f"amepb_create(device={device}) failed ({rc}): {msg}")
        self.handle = handle
        self.device = int(device)

    def error(self) -> str:
        return self.lib.amepb_last_error(self.handle).decode()

    def check(self, rc: int, what: str):
        if rc != 0:
            raise AmepbError(f"{what} failed ({rc}): {self.error()}")

    def kernel_launches(self) -> int:
        return int(self.lib.amepb_kernel_launches(self.handle))

    def rdf_info(self) -> dict:
        info = RdfInfo()
        self.check(self.lib.amepb_rdf_last_info(self.handle, ctypes.byref(info)), "amepb_rdf_last_info")
        return info.as_dict()

    def set_timing(self, enable: bool):
        self.check(self.lib.amepb_set_timing(self.handle, 1 if enable else 0), "amepb_set_timing")

    def last_kernel_ms(self) -> float:
        return float(self.lib.amepb_last_kernel_ms(self.handle))

    def set_symmetric(self, enable: bool):
        self.check(self.lib.amepb_rdf_set_symmetric(self.handle, 1 if enable else 0), "amepb_rdf_set_symmetric")

    def set_shard(self, index: int, count: int):
        """Intra-frame sharding of amepb_rdf_batch (amepb_rdf_set_shard)."""
        self.check(self.lib.amepb_rdf_set_shard(self.handle, int(index), int(count)), "amepb_rdf_set_shard")

    def set_pcf_angle_query_offset(self, offset: int):
        """The queries of amepb_pcf_angle_counts are the slice of the targets from ``offset`` on (< 0: off)."""
        self.check(self.lib.amepb_pcf_angle_set_query_offset(self.handle, int(offset)), "amepb_pcf_angle_set_query_offset")

    def set_pcf2d_query_offset(self, offset: int):
        """The queries of amepb_pcf2d_counts are the slice of the targets from ``offset`` on (< 0: off)."""
        self.check(self.lib.amepb_pcf2d_set_query_offset(self.handle, int(offset)), "amepb_pcf2d_set_query_offset")

    def set_sf2d_shard(self, index: int, count: int):
        """Intra-frame sharding of amepb_sf2d_modes over rows of q tiles (amepb_sf2d_set_shard)."""
        self.check(self.lib.amepb_sf2d_set_shard(self.handle, int(index), int(count)), "amepb_sf2d_set_shard")

    def set_cells_per_cutoff(self, k: int):
        self.check(self.lib.amepb_rdf_set_cells_per_cutoff(self.handle, int(k)),
                   "amepb_rdf_set_cells_per_cutoff")

    def close(self):
        if getattr(self, "handle", None):
            self.lib.amepb_destroy(self.handle)
            self.handle = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass


_tls = threading.local()


def context(device: int = 0) -> Context:
    """Per-thread, per-device context (average_func may call from a thread pool,
    amep/utils.py:174-182; a context is not thread safe)."""
    table = getattr(_tls, "contexts", None)
    if table is None:
        table = _tls.contexts = {}
    ctx = table.get(device)
    if ctx is None:
        ctx = table[device] = Context(device)
    return ctx
