"""The reference's FLAG files, consumed on the GPU.

  RAW  plain uint16 stream (".bin"), the input of `bench decompress -r`
       (benchmark/flagstats.cpp:415-468): read in 1,024,000-byte blocks straight
       into the pinned ring, DMA + kernel overlapped.
  LZ4  the container lz4f()/lz4hc() write (:110-186) and lz4_decompress() reads
       (:288-358): [int32 raw_size][int32 comp_size][LZ4 block] ...  The blocks
       cross PCIe compressed and are decoded on the GPU (one warp per block).
  ZSTD the container zstd() writes (:192-215) and zstd_decompress() reads (:636-676):
       the same records around one Zstandard frame each; decoded on the GPU too.

Thin ctypes front-end of FLAGSTAT_cuda_file_u64 / _container_u64 / _lz4_decode;
no CPU decoder or fallback lives here.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence, Tuple

import numpy as np

from . import _capi
from ._capi import check, lib

RAW, LZ4, ZSTD = 0, 1, 2
SAMTOOLS = 0x100
"""OR into the format: count like FLAGSTAT_cuda_samtools_u64 (flags[0] / flags[16] = the exact
n_pair_all) -- the reference's 'samtools' readers of the same files (flagstats.cpp:496-519, 547-590)."""
_EXT = {".bin": RAW, ".raw": RAW, ".lz4": LZ4, ".zst": ZSTD, ".zstd": ZSTD}


def _format_of(path: str, fmt: Optional[int]) -> int:
    if fmt is not None:
        return int(fmt)
    ext = os.path.splitext(path)[1].lower()
    if ext not in _EXT:
        raise ValueError(f"cannot infer the container format from {path!r}; pass fmt=RAW or fmt=LZ4")
    return _EXT[ext]


def _flags(flags: Optional[np.ndarray]) -> np.ndarray:
    f = np.zeros(32, np.uint64) if flags is None else flags
    if f.dtype != np.uint64 or f.size != 32 or not f.flags["C_CONTIGUOUS"]:
        raise ValueError("flags must be a contiguous uint64[32]")
    return f


def flagstat_file(path: str, fmt: Optional[int] = None, flags: Optional[np.ndarray] = None,
                  samtools: bool = False) -> Tuple[np.ndarray, int]:
    """Accumulate the counters of a FLAG file into ``flags``; returns (flags, n_records)."""
    f = _flags(flags)
    n = C.c_uint64(0)
    check(lib().FLAGSTAT_cuda_file_u64(os.fsencode(path), _format_of(path, fmt) | (SAMTOOLS if samtools else 0),
                                       f.ctypes.data_as(_capi.u64p), C.byref(n)),
          "FLAGSTAT_cuda_file_u64")
    return f, int(n.value)


def flagstat_container(blob, fmt: int, flags: Optional[np.ndarray] = None,
                       samtools: bool = False) -> Tuple[np.ndarray, int]:
    """Same for a container held in host memory (bytes-like or uint8 array)."""
    f = _flags(flags)
    buf = np.frombuffer(blob, dtype=np.uint8) if not isinstance(blob, np.ndarray) else blob
    buf = np.ascontiguousarray(buf, dtype=np.uint8)
    n = C.c_uint64(0)
    check(lib().FLAGSTAT_cuda_container_u64(buf.ctypes.data, buf.size, int(fmt) | (SAMTOOLS if samtools else 0),
                                            f.ctypes.data_as(_capi.u64p), C.byref(n)),
          "FLAGSTAT_cuda_container_u64")
    return f, int(n.value)


def zstd_decode(frames: Sequence[bytes], raw_sizes: Sequence[int]):
    """Decode Zstandard frames on the GPU (FLAGSTAT_cuda_zstd_decode); same return as lz4_decode."""
    return lz4_decode(frames, raw_sizes, _entry="FLAGSTAT_cuda_zstd_decode")


def lz4_decode(blocks: Sequence[bytes], raw_sizes: Sequence[int], _entry: str = "FLAGSTAT_cuda_lz4_decode"):
    """Decode LZ4 blocks on the GPU.  Returns (list of bytes, list of status); status is
    the decoded size, or a negative code for a malformed block (its bytes are undefined)."""
    nb = len(blocks)
    comp_off = np.zeros(nb, np.uint64)
    comp_size = np.array([len(b) for b in blocks], np.uint32)
    raw_size = np.array(list(raw_sizes), np.uint32)
    raw_off = np.zeros(nb, np.uint64)
    c = r = 0
    for i in range(nb):
        c = (c + 15) & ~15
        comp_off[i] = c
        c += int(comp_size[i])
        raw_off[i] = r
        r += (int(raw_size[i]) + 15) & ~15
    comp = np.zeros(max(c, 1), np.uint8)
    for i, b in enumerate(blocks):
        comp[int(comp_off[i]):int(comp_off[i]) + len(b)] = np.frombuffer(b, dtype=np.uint8)
    raw = np.zeros(max(r, 1), np.uint8)
    status = np.zeros(max(nb, 1), np.int32)
    check(getattr(lib(), _entry)(
        comp.ctypes.data, c, comp_off.ctypes.data_as(_capi.u64p), comp_size.ctypes.data_as(_capi.u32p),
        raw_off.ctypes.data_as(_capi.u64p), raw_size.ctypes.data_as(_capi.u32p), nb, raw.ctypes.data, r,
        status.ctypes.data_as(C.POINTER(C.c_int))), _entry)
    out = [raw[int(raw_off[i]):int(raw_off[i]) + int(raw_size[i])].tobytes() for i in range(nb)]
    return out, status[:nb].tolist()


def ingest_text(text, with_flags: bool = False):
    """FLAG text column (one decimal FLAG per line, benchmark/utility.cpp:29-32) -> uint16
    column, parsed on the GPU.  Returns the column, or (column, flags) with the counters of
    that column (counted from device memory) when ``with_flags``."""
    buf = np.frombuffer(text, dtype=np.uint8) if not isinstance(text, np.ndarray) else text
    buf = np.ascontiguousarray(buf, dtype=np.uint8)
    cap = int(np.count_nonzero(buf == 10)) + 1
    out = np.empty(cap, np.uint16)
    n = C.c_uint64(0)
    f = np.zeros(32, np.uint64)
    check(lib().FLAGSTAT_cuda_ingest_text(buf.ctypes.data, buf.size, out.ctypes.data, cap, C.byref(n),
                                          f.ctypes.data_as(_capi.u64p) if with_flags else None),
          "FLAGSTAT_cuda_ingest_text")
    col = out[: int(n.value)].copy()
    return (col, f) if with_flags else col
