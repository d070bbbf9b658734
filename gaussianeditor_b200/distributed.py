"""Multi-GPU host logic for the rasterizer path (one process per GPU, torch.distributed).

The path shards over VIEWS: every rank holds the whole cloud and renders its own camera stream (the units are
independent, so the data path needs no collective -- "weak" scaling, bench.py --gpus N).  When the replicas are
used for data-parallel optimisation the per-Gaussian gradients are summed across ranks with one all-reduce per
parameter group; that is the only exchange step and it lives outside the rasterizer, exactly where a
GaussianEditor batch>1 loop would put it.  The Gaussian-index-sharded path of SURVEY.md 8(e) (BASELINE config 4: one
view, the cloud split over the ranks, tile rows owned round-robin) is a separate module -- sharded.py for the dense
exchange, sparse_sharded.py + csrc/sparse_exchange.cu for the default sparse peer-memory exchange; DESIGN.md 7.2.

Works with the `nccl` backend on GPUs and with `gloo` on CPU tensors (used by the world_size-2 CPU tests).
"""
from __future__ import annotations

from typing import Iterable, List, Sequence

import torch
import torch.distributed as dist


def view_shard(num_views: int, world_size: int, rank: int) -> List[int]:
    """Round-robin assignment of camera indices to ranks; every view is owned by exactly one rank and the
    per-rank counts differ by at most one."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    return list(range(rank, num_views, world_size))


def steps_per_rank(num_views: int, world_size: int) -> int:
    """Number of lock-step iterations needed so that every rank can take part in every collective (ranks with
    one view fewer repeat their last view with zero loss weight)."""
    return (num_views + world_size - 1) // world_size


def allreduce_gradients(tensors: Iterable[torch.Tensor], group=None, average: bool = False) -> None:
    """In-place sum (or mean) of per-Gaussian gradient tensors across ranks. Tensors are flattened into one
    bucket per dtype so a 1M-Gaussian cloud (59 floats per Gaussian) is a single ~236 MB collective."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    tensors = [t for t in tensors if t is not None]
    by_dtype = {}
    for t in tensors:
        by_dtype.setdefault(t.dtype, []).append(t)
    world = dist.get_world_size(group)
    for _, ts in by_dtype.items():
        flat = torch.cat([t.reshape(-1) for t in ts])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        if average:
            flat /= world
        off = 0
        for t in ts:
            n = t.numel()
            t.copy_(flat[off:off + n].view_as(t))
            off += n


def max_over_ranks(value: float, device=None, group=None) -> float:
    """Device-time reduction used for every multi-GPU timing (bench.py): the job is as slow as its slowest rank."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t[0])
