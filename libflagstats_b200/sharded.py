"""Range-sharded flagstat over the GPUs of one box: one process per GPU
(torch.distributed, NCCL over NVLink/NVSwitch), the FLAG column split into
contiguous ranges, the single-GPU kernel per shard, and ONE exchange: an
all-reduce of the 32 64-bit counters (256 bytes).

The counters are a plain sum over records (f(A||B) = f(A) + f(B), the same
property that lets the reference reuse one counters[32] across blocks,
benchmark/flagstats.cpp:304,329), so no other communication exists on the path.
The reference has no multi-device code at all.

Two implementations of the exchange:

  allreduce_counters / flagstat_sharded   kernel, then a separate NCCL all-reduce
                                          (gloo on CPU tensors: the CPU tests)
  FusedExchange / flagstat_sharded_fused  ONE kernel per rank that counts the
                                          shard and exchanges the 32 counters
                                          through peer-mapped memory over NVLink
                                          (FLAGSTAT_cuda_device_allreduce); the
                                          process group is only used once, to
                                          pass the 64-byte IPC handles around.
"""
from __future__ import annotations

import ctypes as C
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


def flagstat_sharded(local_values, out=None, group=None, stream=None, samtools: bool = False):
    """Each rank passes its own shard (a CUDA tensor of 16-bit FLAG words already
    resident in its GPU's HBM).  Returns the int64[32] CUDA tensor holding the
    GLOBAL counters on every rank.  Asynchronous with respect to the host."""
    from . import flagstat_device

    out = flagstat_device(local_values, out=out, stream=stream, samtools=samtools)
    return allreduce_counters(out, group=group)


class FusedExchange:
    """Per-rank handle of the fused count + exchange path
    (include/flagstats_cuda.h, FLAGSTAT_cuda_xchg_*).

    One instance per process/GPU.  Construction is a collective over `group`:
    every rank allocates its exchange buffer on its current CUDA device, the
    CUDA IPC handles are all-gathered through torch.distributed (any backend)
    and every rank maps every peer's buffer.  After that no host-side
    communication happens on the path.
    """

    def __init__(self, group=None, device=None, timeout_ms: int = 20000, overlap: bool = False):
        import torch
        import torch.distributed as dist

        from ._capi import check, lib

        self._lib = lib()
        self._check = check
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device())
        self.device = torch.device(device)
        have_pg = dist.is_available() and dist.is_initialized()
        self.world = dist.get_world_size(group) if have_pg else 1
        self.rank = dist.get_rank(group) if have_pg else 0
        self._h = C.c_void_p()
        self._pending_out = None
        handle = (C.c_ubyte * 64)()
        with torch.cuda.device(self.device):
            check(self._lib.FLAGSTAT_cuda_xchg_create(C.byref(self._h), self.rank, self.world,
                                                      C.cast(handle, C.c_void_p)),
                  "FLAGSTAT_cuda_xchg_create")
            if self.world > 1:
                mine = torch.tensor(list(bytes(handle)), dtype=torch.uint8)
                if dist.get_backend(group) == "nccl":
                    mine = mine.to(self.device)
                gathered = [torch.empty_like(mine) for _ in range(self.world)]
                dist.all_gather(gathered, mine, group=group)
                blob = b"".join(bytes(t.cpu().tolist()) for t in gathered)
                buf = C.create_string_buffer(blob, len(blob))
                check(self._lib.FLAGSTAT_cuda_xchg_connect(self._h, C.cast(buf, C.c_void_p)),
                      "FLAGSTAT_cuda_xchg_connect")
                dist.barrier(group=group)
            check(self._lib.FLAGSTAT_cuda_xchg_set_timeout_ms(self._h, int(timeout_ms)),
                  "FLAGSTAT_cuda_xchg_set_timeout_ms")
        self.set_overlap(overlap)

    def set_overlap(self, on: bool) -> bool:
        """Overlapped steps (FLAGSTAT_cuda_xchg_set_overlap): back-to-back calls on one stream
        may start before the previous one has finished exchanging; the caller guarantees the
        shard is not written by the kernel enqueued immediately before a call."""
        return bool(self._lib.FLAGSTAT_cuda_xchg_set_overlap(self._h, 1 if on else 0))

    def flagstat(self, local_values, out=None, accumulate: bool = False, stream=None,
                 pospopcnt: bool = False, samtools: bool = False, deferred: bool = False):
        """Count this rank's shard and return the GLOBAL counters in `out`
        (int64[32] CUDA tensor, or [16] for pospopcnt) on every rank.  One
        kernel launch; asynchronous with respect to the host.  Collective.
        ``samtools``: also the exact n_pair_all in slots 0 / 16
        (FLAGSTAT_cuda_samtools_device_allreduce).
        ``deferred`` (flagstat mode): FLAGSTAT_cuda_device_allreduce_deferred -- the launch
        pushes this rank's totals but does not wait for the peers'; `out` is written by the
        next ``flagstat`` call on this handle or by ``collect()`` (keep it alive until then)."""
        import torch

        from . import _device_view, _stream_ptr

        ptr, n, _keep = _device_view(local_values)
        nout = 16 if pospopcnt else 32
        if out is None:
            out = torch.zeros(nout, dtype=torch.int64, device=local_values.device)
        if out.dtype != torch.int64 or out.numel() != nout or not out.is_cuda:
            raise ValueError(f"out must be a CUDA int64[{nout}] tensor")
        if stream is None:
            stream = torch.cuda.current_stream(local_values.device)
        if pospopcnt and samtools:
            raise ValueError("pospopcnt and samtools are different modes")
        if deferred and (pospopcnt or samtools):
            raise ValueError("deferred collection is available for the flagstat mode only")
        fn = (self._lib.POSPOPCNT_cuda_device_allreduce if pospopcnt
              else self._lib.FLAGSTAT_cuda_samtools_device_allreduce if samtools
              else self._lib.FLAGSTAT_cuda_device_allreduce_deferred if deferred
              else self._lib.FLAGSTAT_cuda_device_allreduce)
        if deferred:
            self._pending_out = out  # keeps the tensor alive until it has been written
        with torch.cuda.device(local_values.device):
            self._check(fn(self._h, ptr, n, out.data_ptr(), 1 if accumulate else 0,
                           _stream_ptr(stream)), "FLAGSTAT_cuda_device_allreduce")
        return out

    def collect(self, stream=None) -> None:
        """FLAGSTAT_cuda_xchg_collect: write the counters of the last deferred call (a one-warp
        kernel on `stream`, which must be ordered after that call).  Collective."""
        import torch

        if stream is None:
            stream = torch.cuda.current_stream(self.device)
        from . import _stream_ptr

        with torch.cuda.device(self.device):
            self._check(self._lib.FLAGSTAT_cuda_xchg_collect(self._h, _stream_ptr(stream)),
                        "FLAGSTAT_cuda_xchg_collect")

    def status(self) -> None:
        """Synchronise and raise if a peer never delivered its counters."""
        self._check(self._lib.FLAGSTAT_cuda_xchg_status(self._h), "FLAGSTAT_cuda_xchg_status")

    def close(self) -> None:
        if self._h:
            self._lib.FLAGSTAT_cuda_xchg_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def flagstat_sharded_fused(local_values, xchg: FusedExchange, out=None, accumulate: bool = False,
                           stream=None):
    """flagstat_sharded with the exchange fused into the counting kernel."""
    return xchg.flagstat(local_values, out=out, accumulate=accumulate, stream=stream)
