"""libflagstats_b200 -- B200-native flagstat hot path behind the reference's API.

Host-side mirror of the reference's Python surface (python/libflagstats.pyx:8-37
in mklarqvist/libflagstats): ``flagstats(values)`` has the same name, argument
meaning, error behaviour and result dict as ``pyflagstats.flagstats``; the
counting itself runs in hand-written sm_100a CUDA behind the C ABI declared in
include/flagstats_cuda.h.  There is no CPU fallback: if libflagstats_cuda.so is
missing or no GPU is present, calls raise.

Beyond the reference surface:
  flagstat_u64 / flagstat_u32   counters as arrays (the FLAGSTAT_* contract)
  pospopcnt_u16                 raw 16-counter mode (STORM_pospopcnt_u16)
  flagstat_samtools_u64 / samtools_stats / samtools_text
                                the reference benchmark's flagstat_loop caller:
                                bam_flagstat_t with the exact n_pair_all and its report
  flagstat_device               async, device-resident (torch CUDA tensors or
                                anything with __cuda_array_interface__)
  BlockStream                   pinned ring for 1,024,000-byte block streaming
  sharded                       range-sharded multi-GPU + all-reduce
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import _capi
from ._capi import FlagstatCudaError, check, lib  # noqa: F401

__version__ = "0.1.0"

SAM_FLAG_NAMES = ["FPAIRED", "FPROPER_PAIR", "FUNMAP", "FMUNMAP", "FREVERSE", "FMREVERSE",
                  "FREAD1", "FREAD2", "FSECONDARY", "FQCFAIL", "FDUP", "FSUPPLEMENTARY",
                  "n_pair_good", "n_sgltn", "n_pair_map"]  # python/libflagstats.pyx:24

CORE20 = (2, 6, 7, 8, 9, 10, 11, 12, 13, 14, 18, 22, 23, 24, 25, 26, 27, 28, 29, 30)
"""Slots FLAGSTAT_cuda writes (CORE19 + slot 9), see include/flagstats_cuda.h."""

BLOCK_RECORDS = 512_000
"""Records per reference block: 1,024,000 bytes (benchmark/flagstats.cpp:119)."""


# --------------------------------------------------------------------------
# pointer plumbing
# --------------------------------------------------------------------------
def _is_torch(x) -> bool:
    return type(x).__module__.split(".")[0] == "torch" and hasattr(x, "data_ptr")


def _device_view(x):
    """(ptr, n_records, keepalive) for a device array of 16-bit elements."""
    if _is_torch(x):
        if not x.is_cuda:
            raise ValueError("expected a CUDA tensor")
        if x.element_size() != 2 or x.dtype.is_floating_point:
            raise ValueError("Values must have a 16-bit integer dtype")
        if not x.is_contiguous():
            raise ValueError("device input must be contiguous")
        return x.data_ptr(), x.numel(), x
    cai = getattr(x, "__cuda_array_interface__", None)
    if cai is None:
        raise ValueError("expected a torch CUDA tensor or an object with __cuda_array_interface__")
    if cai["typestr"] not in ("<u2", "<i2", "|u2", "=u2"):
        raise ValueError("Values must have the dtype \"uint16\"")
    if cai.get("strides") is not None:
        raise ValueError("device input must be contiguous")
    n = 1
    for d in cai["shape"]:
        n *= int(d)
    return int(cai["data"][0]), n, x


def _is_device_array(x) -> bool:
    return (_is_torch(x) and x.is_cuda) or hasattr(x, "__cuda_array_interface__")


def _host_u16(values) -> np.ndarray:
    # error behaviour of python/libflagstats.pyx:9-17
    if type(values) != np.ndarray:  # noqa: E721  (the reference tests the exact type)
        raise ValueError("Values must be an numpy.ndarray")
    if values.dtype != "uint16":
        raise ValueError("Values must have the dtype \"uint16\"")
    if not values.flags["C_CONTIGUOUS"]:
        print("Input array is not contiguous. Fixing...")
        values = np.ascontiguousarray(values, dtype=np.uint16)
    return values


def _stream_ptr(stream) -> int:
    if stream is None:
        return 0
    if isinstance(stream, int):
        return stream
    return int(stream.cuda_stream)  # torch.cuda.Stream


# --------------------------------------------------------------------------
# the FLAGSTAT_* contract
# --------------------------------------------------------------------------
def flagstat_u64(values, flags: Optional[np.ndarray] = None) -> np.ndarray:
    """FLAGSTAT_cuda_u64: accumulate the 32 counters of ``values`` into
    ``flags`` (uint64[32], zeros if omitted).  Host numpy or device arrays."""
    f = np.zeros(32, np.uint64) if flags is None else flags
    if f.dtype != np.uint64 or f.size != 32 or not f.flags["C_CONTIGUOUS"]:
        raise ValueError("flags must be a contiguous uint64[32]")
    if _is_device_array(values):
        ptr, n, _keep = _device_view(values)
    else:
        values = _host_u16(values)
        ptr, n = values.ctypes.data, values.size
    check(lib().FLAGSTAT_cuda_u64(ptr, n, f.ctypes.data_as(_capi.u64p)), "FLAGSTAT_cuda_u64")
    return f


def flagstat_u32(values, flags: Optional[np.ndarray] = None) -> np.ndarray:
    """FLAGSTAT_cuda with the reference's own signature
    ``int f(const uint16_t*, uint32_t len, uint32_t* flags)``
    (libflagstats.h:2970): uint32 counters, accumulate, wrap mod 2^32."""
    f = np.zeros(32, np.uint32) if flags is None else flags
    if f.dtype != np.uint32 or f.size != 32 or not f.flags["C_CONTIGUOUS"]:
        raise ValueError("flags must be a contiguous uint32[32]")
    if _is_device_array(values):
        ptr, n, _keep = _device_view(values)
    else:
        values = _host_u16(values)
        ptr, n = values.ctypes.data, values.size
    if n > 0xFFFFFFFF:
        raise ValueError("len does not fit the reference's uint32_t; use flagstat_u64")
    check(lib().FLAGSTAT_cuda(ptr, n, f.ctypes.data_as(_capi.u32p)), "FLAGSTAT_cuda")
    return f


def pospopcnt_u16(values) -> np.ndarray:
    """STORM_pospopcnt_u16 (libalgebra.h:3496): out[j] = #records with bit j
    set; returned as uint64[16] (the reference's uint32 out wraps at 2^32)."""
    out = np.zeros(16, np.uint64)
    if _is_device_array(values):
        ptr, n, _keep = _device_view(values)
    else:
        values = _host_u16(values)
        ptr, n = values.ctypes.data, values.size
    check(lib().POSPOPCNT_cuda_u16_u64(ptr, n, out.ctypes.data_as(_capi.u64p)),
          "POSPOPCNT_cuda_u16_u64")
    return out


def flagstat_device(values, out=None, stream=None, pospopcnt: bool = False, samtools: bool = False,
                    overlap: bool = False):
    """Asynchronous, device-resident form (FLAGSTAT_cuda_device /
    POSPOPCNT_cuda_device / FLAGSTAT_cuda_samtools_device).  ``values``: torch CUDA
    tensor of 16-bit integers.  ``out``: torch.int64[32] (or [16]) CUDA tensor that is
    ACCUMULATED into; allocated zeroed if omitted.  Returns ``out`` without
    synchronising.  ``samtools``: also the exact n_pair_all in slots 0 / 16.
    ``overlap`` (flagstat mode): FLAGSTAT_cuda_device_overlapped -- the kernel may start
    before the previous kernel of the stream has finished; ``values`` must not be written
    by that kernel."""
    import torch

    ptr, n, _keep = _device_view(values)
    nout = 16 if pospopcnt else 32
    if out is None:
        out = torch.zeros(nout, dtype=torch.int64, device=values.device)
    if out.dtype != torch.int64 or out.numel() != nout or not out.is_cuda:
        raise ValueError(f"out must be a CUDA int64[{nout}] tensor")
    if stream is None:
        stream = torch.cuda.current_stream(values.device)
    if pospopcnt and samtools:
        raise ValueError("pospopcnt and samtools are different modes")
    if overlap and (pospopcnt or samtools):
        raise ValueError("overlap is available for the flagstat mode only")
    fn = (lib().POSPOPCNT_cuda_device if pospopcnt
          else lib().FLAGSTAT_cuda_samtools_device if samtools
          else lib().FLAGSTAT_cuda_device_overlapped if overlap else lib().FLAGSTAT_cuda_device)
    with torch.cuda.device(values.device):
        check(fn(ptr, n, out.data_ptr(), _stream_ptr(stream)), "FLAGSTAT_cuda_device")
    return out


def counters_to_dict(flags, n_values: int) -> dict:
    """The result dict of python/libflagstats.pyx:26-35."""
    flags = np.asarray(flags)
    ret = {
        "n_values": n_values,
        "passed": dict(zip(SAM_FLAG_NAMES, flags[0:15, ])),
        "failed": dict(zip(SAM_FLAG_NAMES, flags[16:31, ])),
    }
    ret["passed"]["mapped"] = n_values - ret["passed"]["FUNMAP"] - ret["failed"]["FUNMAP"]
    ret["passed"]["paired_in_seq"] = ret["passed"]["FREAD1"] + ret["passed"]["FREAD2"]
    return ret


def flagstats(values) -> dict:
    """Drop-in for ``pyflagstats.flagstats`` (python/libflagstats.pyx:8-37).

    Same checks, same dict.  Counters are uint32 like the reference's; device
    arrays are accepted in addition to numpy arrays."""
    if _is_device_array(values):
        n = _device_view(values)[1]
        flags = flagstat_u32(values)
    else:
        values = _host_u16(values)
        n = len(values)
        flags = flagstat_u32(values)
    return counters_to_dict(flags, n)


SAMTOOLS_FIELDS = ("n_reads", "n_mapped", "n_pair_all", "n_pair_map", "n_pair_good", "n_sgltn",
                   "n_read1", "n_read2", "n_dup", "n_diffchr", "n_diffhigh", "n_secondary", "n_supp")
"""bam_flagstat_t in declaration order (benchmark/flagstats.cpp:43-49); each [QC-pass, QC-fail]."""


def flagstat_samtools_u64(values, flags: Optional[np.ndarray] = None) -> np.ndarray:
    """FLAGSTAT_cuda_samtools_u64: flagstat_u64 plus the exact 'paired in sequencing' count
    (n_pair_all, benchmark/flagstats.cpp:58-59) in slots 0 / 16.  Accumulates."""
    f = np.zeros(32, np.uint64) if flags is None else flags
    if f.dtype != np.uint64 or f.size != 32 or not f.flags["C_CONTIGUOUS"]:
        raise ValueError("flags must be a contiguous uint64[32]")
    if _is_device_array(values):
        ptr, n, _keep = _device_view(values)
    else:
        values = _host_u16(values)
        ptr, n = values.ctypes.data, values.size
    check(lib().FLAGSTAT_cuda_samtools_u64(ptr, n, f.ctypes.data_as(_capi.u64p)),
          "FLAGSTAT_cuda_samtools_u64")
    return f


def samtools_stats(values, stats: Optional[np.ndarray] = None) -> np.ndarray:
    """FLAGSTAT_cuda_samtools: the reference benchmark's flagstat_loop
    (benchmark/flagstats.cpp:51-71) over a host or device column.  Returns (and adds to)
    a bam_flagstat_t as int64[13, 2], rows in SAMTOOLS_FIELDS order."""
    s = np.zeros((13, 2), np.int64) if stats is None else stats
    if s.dtype != np.int64 or s.shape != (13, 2) or not s.flags["C_CONTIGUOUS"]:
        raise ValueError("stats must be a contiguous int64[13, 2]")
    if _is_device_array(values):
        ptr, n, _keep = _device_view(values)
    else:
        values = _host_u16(values)
        ptr, n = values.ctypes.data, values.size
    check(lib().FLAGSTAT_cuda_samtools(ptr, n, s.ctypes.data), "FLAGSTAT_cuda_samtools")
    return s


def samtools_stats_from_counters(flags) -> np.ndarray:
    """bam_flagstat_t (int64[13, 2]) from 32 counters of a *_samtools entry."""
    f = np.ascontiguousarray(np.asarray(flags), dtype=np.uint64)
    if f.size != 32:
        raise ValueError("flags must hold 32 counters")
    s = np.zeros((13, 2), np.int64)
    check(lib().FLAGSTAT_cuda_samtools_from_counters(f.ctypes.data_as(_capi.u64p), s.ctypes.data),
          "FLAGSTAT_cuda_samtools_from_counters")
    return s


def samtools_text(stats) -> str:
    """FLAGSTAT_cuda_samtools_report: the report of benchmark/flagstats.cpp:577-588, byte
    for byte, from a bam_flagstat_t (int64[13, 2])."""
    s = np.ascontiguousarray(np.asarray(stats), dtype=np.int64)
    if s.size != 26:
        raise ValueError("stats must be an int64[13, 2] bam_flagstat_t")
    buf = C.create_string_buffer(2048)
    n = lib().FLAGSTAT_cuda_samtools_report(s.ctypes.data, buf, len(buf))
    if n < 0:
        raise FlagstatCudaError(n, "FLAGSTAT_cuda_samtools_report")
    return buf.value.decode()


def samtools_report(flags) -> str:
    """Text report in samtools-flagstat order from the 32 counters of the plain
    FLAGSTAT_* contract (cf. benchmark/flagstats.cpp:577-588).  'paired in sequencing'
    uses the wrapper's READ1+READ2 approximation (python/libflagstats.pyx:35); use
    samtools_stats + samtools_text for the exact count and the reference's own
    formatting."""
    f = [int(x) for x in np.asarray(flags)]
    p, q = f[:16], f[16:]

    def pct(a, b):
        return "N/A" if b == 0 else f"{100.0 * a / b:.2f}%"

    rows = [
        (p[9], q[9], "in total (QC-passed reads + QC-failed reads)"),
        (p[8], q[8], "secondary"),
        (p[11], q[11], "supplementary"),
        (p[10], q[10], "duplicates"),
        (p[9] - p[2], q[9] - q[2], f"mapped ({pct(p[9] - p[2], p[9])} : {pct(q[9] - q[2], q[9])})"),
        (p[6] + p[7], q[6] + q[7], "paired in sequencing"),
        (p[6], q[6], "read1"),
        (p[7], q[7], "read2"),
        (p[12], q[12], f"properly paired ({pct(p[12], p[6] + p[7])} : {pct(q[12], q[6] + q[7])})"),
        (p[14], q[14], "with itself and mate mapped"),
        (p[13], q[13], f"singletons ({pct(p[13], p[6] + p[7])} : {pct(q[13], q[6] + q[7])})"),
    ]
    return "\n".join(f"{a} + {b} {t}" for a, b, t in rows)


# --------------------------------------------------------------------------
# block streaming (benchmark/flagstats.cpp:288-358 caller pattern)
# --------------------------------------------------------------------------
class BlockStream:
    """Pinned ring of ``n_slots`` blocks; H2D of block k+1 overlaps the kernel
    of block k on separate CUDA streams; one device-side counter set is shared
    by all blocks (like the reference's shared counters[32], flagstats.cpp:304)."""

    DMA, ZEROCOPY = 0, 1

    def __init__(self, device: int = 0, block_records: int = BLOCK_RECORDS, n_slots: int = 4,
                 mode: Optional[int] = None, coalesce: int = 0):
        self._h = C.c_void_p()
        self.block_records = int(block_records)
        if mode is None:
            check(lib().FLAGSTAT_cuda_stream_open(C.byref(self._h), device, self.block_records,
                                                  n_slots), "FLAGSTAT_cuda_stream_open")
        else:
            check(lib().FLAGSTAT_cuda_stream_open_ex(C.byref(self._h), device, self.block_records,
                                                     n_slots, int(mode), int(coalesce)),
                  "FLAGSTAT_cuda_stream_open_ex")

    def acquire(self) -> np.ndarray:
        """Next pinned slot as a writable uint16 view (zero-copy producer)."""
        p = lib().FLAGSTAT_cuda_stream_acquire(self._h)
        if not p:
            raise FlagstatCudaError(_capi.lib().FLAGSTAT_cuda_available() and -4 or -1,
                                    "FLAGSTAT_cuda_stream_acquire")
        buf = (C.c_uint16 * self.block_records).from_address(p)
        return np.frombuffer(buf, dtype=np.uint16)

    def submit(self, n_records: int) -> None:
        check(lib().FLAGSTAT_cuda_stream_submit(self._h, int(n_records)),
              "FLAGSTAT_cuda_stream_submit")

    def push(self, block: np.ndarray) -> None:
        block = _host_u16(block)
        check(lib().FLAGSTAT_cuda_stream_push(self._h, block.ctypes.data, block.size),
              "FLAGSTAT_cuda_stream_push")

    def finish(self, flags: Optional[np.ndarray] = None) -> np.ndarray:
        f = np.zeros(32, np.uint64) if flags is None else flags
        check(lib().FLAGSTAT_cuda_stream_finish(self._h, f.ctypes.data_as(_capi.u64p)),
              "FLAGSTAT_cuda_stream_finish")
        return f

    def selftime(self, n_blocks: int):
        """(counters, seconds) of re-submitting the ring's contents as n_blocks full
        blocks from C (FLAGSTAT_cuda_stream_selftime)."""
        f = np.zeros(32, np.uint64)
        sec = C.c_double()
        check(lib().FLAGSTAT_cuda_stream_selftime(self._h, int(n_blocks), f.ctypes.data_as(_capi.u64p),
                                                  C.byref(sec)), "FLAGSTAT_cuda_stream_selftime")
        return f, sec.value

    def close(self) -> None:
        if self._h:
            lib().FLAGSTAT_cuda_stream_close(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def available() -> int:
    """Number of usable CUDA devices (0 = none)."""
    return int(lib().FLAGSTAT_cuda_available())
