"""Frame sharding over ranks and the single all-reduce that follows.

Frames are independent units of the path (``average_func`` maps a pure
per-frame function, amep/utils.py:156-187), so each rank (one process per GPU,
``torch.distributed`` over NCCL/NVLink) evaluates a contiguous block of the
selected frames, writes its rows into a zero-initialised table and one
``all_reduce(SUM)`` assembles the table everywhere.  Rows are integers
(pair counts) or float64 sums added to zeros, so the result is independent of
the number of ranks, bit for bit.
"""
from __future__ import annotations

import numpy as np


def world(group=None):
    """-> (rank, world_size); (0, 1) when torch.distributed is not initialised."""
    try:
        import torch.distributed as dist
    except Exception:  # pragma: no cover
        return 0, 1
    if not dist.is_available() or not dist.is_initialized():
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


def shard_bounds(nitems: int, rank: int, world_size: int):
    """Contiguous block ``[lo, hi)`` of ``nitems`` for ``rank``; sizes differ by at most one."""
    base, extra = divmod(nitems, world_size)
    lo = rank * base + min(rank, extra)
    hi = lo + base + (1 if rank < extra else 0)
    return lo, hi


def assemble_rows(local_rows, lo: int, total_rows: int, group=None):
    """Place ``local_rows`` (tensor or ndarray ``[hi-lo, ...]``) at ``[lo:hi]`` of
    a zero table ``[total_rows, ...]`` and all-reduce it (SUM) over ``group``.
    Returns a NumPy array.  With a single rank nothing is communicated."""
    import torch
    import torch.distributed as dist
    rank, ws = world(group)
    is_tensor = isinstance(local_rows, torch.Tensor)
    if ws == 1:
        return local_rows.cpu().numpy() if is_tensor else np.asarray(local_rows)
    t = local_rows if is_tensor else torch.from_numpy(np.ascontiguousarray(local_rows))
    backend = dist.get_backend(group)
    if backend == "nccl" and not t.is_cuda:
        t = t.cuda()
    table = torch.zeros((total_rows,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    table[lo:lo + t.shape[0]] = t
    if table.is_complex():
        dist.all_reduce(torch.view_as_real(table), op=dist.ReduceOp.SUM, group=group)
    else:
        dist.all_reduce(table, op=dist.ReduceOp.SUM, group=group)
    return table.cpu().numpy()


def _to_backend_tensor(array, group):
    import torch
    import torch.distributed as dist
    t = array if isinstance(array, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(array))
    if dist.get_backend(group) == "nccl" and not t.is_cuda:
        t = t.cuda()
    return t


def all_reduce_sum(array, group=None):
    """Element-wise sum of a NumPy array / tensor over the ranks -> NumPy array (one all-reduce)."""
    import torch
    import torch.distributed as dist
    t = _to_backend_tensor(array, group).clone()
    if t.is_complex():
        dist.all_reduce(torch.view_as_real(t), op=dist.ReduceOp.SUM, group=group)
    else:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.cpu().numpy()


def agree_shape(shape, group=None, max_dims: int = 8):
    """The one shape all ranks that have one (``shape`` not None) share; ranks without work pass
    None and learn it.  Decided by two MAX all-reduces, so that a rank with an empty shard neither
    raises on its own nor leaves the others waiting in a collective."""
    import torch
    import torch.distributed as dist
    pad = torch.zeros(max_dims + 1, dtype=torch.int64)
    if shape is not None:
        if len(shape) > max_dims:
            raise ValueError("agree_shape: too many dimensions")
        pad[0] = len(shape)
        for i, s in enumerate(shape):
            pad[1 + i] = int(s)
    pad = _to_backend_tensor(pad, group)
    dist.all_reduce(pad, op=dist.ReduceOp.MAX, group=group)
    pad = pad.cpu().tolist()
    return tuple(int(x) for x in pad[1:1 + int(pad[0])])


def all_true(flag: bool, group=None) -> bool:
    """Logical AND of ``flag`` over the ranks."""
    import torch
    import torch.distributed as dist
    t = _to_backend_tensor(torch.tensor([0 if flag else 1], dtype=torch.int64), group)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return int(t.item()) == 0
