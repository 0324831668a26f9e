"""Range-sharded flagstat over the GPUs of one box: one process per GPU
(torch.distributed, NCCL over NVLink/NVSwitch), the FLAG column split into
contiguous ranges, the single-GPU kernel per shard, and ONE exchange: an
all-reduce of the 32 64-bit counters (256 bytes).

The counters are a plain sum over records (f(A||B) = f(A) + f(B), the same
property that lets the reference reuse one counters[32] across blocks,
benchmark/flagstats.cpp:304,329), so no other communication exists on the path.
The reference has no multi-device code at all.
"""
from __future__ import annotations

from typing import Tuple

SHARD_ALIGN = 8  # records; keeps every shard base 16-byte aligned


def shard_range(n: int, world: int, rank: int, align: int = SHARD_ALIGN) -> Tuple[int, int]:
    """Contiguous range [lo, hi) of rank `rank` out of `world`; interior
    boundaries are multiples of `align` records, the last shard takes the
    ragged end.  The ranges tile [0, n) exactly."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad world/rank")
    per = n // world

    def edge(r: int) -> int:
        return n if r == world else (r * per) // align * align

    return edge(rank), edge(rank + 1)


def allreduce_counters(counters, group=None):
    """In-place SUM all-reduce of an int64 counter tensor across ranks.
    NCCL for CUDA tensors, gloo for CPU tensors.  No-op without a process
    group.  Counters are exact integers, so the result is order-independent."""
    import torch
    import torch.distributed as dist

    if counters.dtype != torch.int64:
        raise ValueError("counters must be int64 (uint64 bit patterns; sums wrap identically)")
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(counters, op=dist.ReduceOp.SUM, group=group)
    return counters


def flagstat_sharded(local_values, out=None, group=None, stream=None):
    """Each rank passes its own shard (a CUDA tensor of 16-bit FLAG words already
    resident in its GPU's HBM).  Returns the int64[32] CUDA tensor holding the
    GLOBAL counters on every rank.  Asynchronous with respect to the host."""
    from . import flagstat_device

    out = flagstat_device(local_values, out=out, stream=stream)
    return allreduce_counters(out, group=group)
