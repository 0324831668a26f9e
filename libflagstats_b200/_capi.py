"""ctypes binding of libflagstats_cuda.so (include/flagstats_cuda.h).

The library is the product; this module only loads it and declares
signatures.  It fails loudly if the shared object is missing -- there is no
CPU fallback anywhere in this package.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# LIBFLAGSTATS_CUDA_SO: load another build of the same ABI instead (the -DFSB_ALL_VARIANTS A/B
# build in tools/bin/, for the variant tests and tools); the product path is the in-tree library
SO_PATH = os.environ.get("LIBFLAGSTATS_CUDA_SO") or os.path.join(HERE, "libflagstats_cuda.so")

u16p = C.POINTER(C.c_uint16)
u32p = C.POINTER(C.c_uint32)
u64p = C.POINTER(C.c_uint64)

# every symbol include/flagstats_cuda.h declares: (restype, argtypes)
SIGNATURES = {
    "FLAGSTAT_cuda": (C.c_int, [C.c_void_p, C.c_uint32, u32p]),
    "FLAGSTAT_cuda_u64": (C.c_int, [C.c_void_p, C.c_uint64, u64p]),
    "FLAGSTAT_cuda_device": (C.c_int, [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]),
    "FLAGSTAT_cuda_device_overlapped": (C.c_int, [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]),
    "FLAGSTAT_cuda_available": (C.c_int, []),
    "FLAGSTAT_cuda_min_len": (C.c_uint32, []),
    "FLAGSTAT_cuda_set_min_len": (None, [C.c_uint32]),
    "FLAGSTAT_cuda_samtools_u64": (C.c_int, [C.c_void_p, C.c_uint64, u64p]),
    "FLAGSTAT_cuda_samtools_device": (C.c_int, [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]),
    "FLAGSTAT_cuda_samtools": (C.c_int, [C.c_void_p, C.c_uint64, C.c_void_p]),
    "FLAGSTAT_cuda_samtools_from_counters": (C.c_int, [u64p, C.c_void_p]),
    "FLAGSTAT_cuda_samtools_report": (C.c_int, [C.c_void_p, C.c_char_p, C.c_size_t]),
    "FLAGSTAT_cuda_samtools_device_allreduce": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p,
                                                          C.c_int, C.c_void_p]),
    "POSPOPCNT_cuda_u16": (C.c_int, [C.c_void_p, C.c_size_t, u32p]),
    "POSPOPCNT_cuda_u16_u64": (C.c_int, [C.c_void_p, C.c_uint64, u64p]),
    "POSPOPCNT_cuda_device": (C.c_int, [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]),
    "FLAGSTAT_cuda_stream_open": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_uint32, C.c_int]),
    "FLAGSTAT_cuda_stream_open_ex": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_uint32, C.c_int,
                                               C.c_int, C.c_int]),
    "FLAGSTAT_cuda_stream_acquire": (C.c_void_p, [C.c_void_p]),
    "FLAGSTAT_cuda_stream_submit": (C.c_int, [C.c_void_p, C.c_uint32]),
    "FLAGSTAT_cuda_stream_push": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32]),
    "FLAGSTAT_cuda_stream_finish": (C.c_int, [C.c_void_p, u64p]),
    "FLAGSTAT_cuda_stream_close": (C.c_int, [C.c_void_p]),
    "FLAGSTAT_cuda_stream_selftime": (C.c_int, [C.c_void_p, C.c_uint32, u64p, C.POINTER(C.c_double)]),
    "FLAGSTAT_cuda_file_u64": (C.c_int, [C.c_char_p, C.c_int, u64p, u64p]),
    "FLAGSTAT_cuda_container_u64": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int, u64p, u64p]),
    "FLAGSTAT_cuda_lz4_decode": (C.c_int, [C.c_void_p, C.c_uint64, u64p, u32p, u64p, u32p, C.c_uint32,
                                           C.c_void_p, C.c_uint64, C.POINTER(C.c_int)]),
    "FLAGSTAT_cuda_zstd_decode": (C.c_int, [C.c_void_p, C.c_uint64, u64p, u32p, u64p, u32p, C.c_uint32,
                                            C.c_void_p, C.c_uint64, C.POINTER(C.c_int)]),
    "FLAGSTAT_cuda_ingest_text": (C.c_int, [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, u64p, u64p]),
    "FLAGSTAT_cuda_multi_u64": (C.c_int, [C.c_void_p, C.c_uint64, u64p, C.c_int]),
    "FLAGSTAT_cuda_xchg_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_void_p]),
    "FLAGSTAT_cuda_xchg_connect": (C.c_int, [C.c_void_p, C.c_void_p]),
    "FLAGSTAT_cuda_xchg_connect_local": (C.c_int, [C.POINTER(C.c_void_p), C.c_int]),
    "FLAGSTAT_cuda_device_allreduce": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p,
                                                 C.c_int, C.c_void_p]),
    "FLAGSTAT_cuda_device_allreduce_deferred": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p,
                                                          C.c_int, C.c_void_p]),
    "FLAGSTAT_cuda_xchg_collect": (C.c_int, [C.c_void_p, C.c_void_p]),
    "POSPOPCNT_cuda_device_allreduce": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p,
                                                  C.c_int, C.c_void_p]),
    "FLAGSTAT_cuda_xchg_set_overlap": (C.c_int, [C.c_void_p, C.c_int]),
    "FLAGSTAT_cuda_xchg_set_timeout_ms": (C.c_int, [C.c_void_p, C.c_uint32]),
    "FLAGSTAT_cuda_xchg_status": (C.c_int, [C.c_void_p]),
    "FLAGSTAT_cuda_xchg_destroy": (C.c_int, [C.c_void_p]),
    "FLAGSTAT_cuda_strerror": (C.c_char_p, [C.c_int]),
    "FLAGSTAT_cuda_version": (C.c_char_p, []),
    "FLAGSTAT_cuda_launch_count": (C.c_uint64, []),
    "FLAGSTAT_cuda_set_variant": (C.c_int, [C.c_int]),
    "FLAGSTAT_cuda_kernel_name": (C.c_char_p, [C.c_int]),
    "FLAGSTAT_cuda_set_lz4_variant": (C.c_int, [C.c_int]),
    "FLAGSTAT_cuda_set_ctas_per_sm": (C.c_int, [C.c_int]),
    "FLAGSTAT_cuda_set_dynamic": (C.c_int, [C.c_longlong, C.c_int]),
    "FLAGSTAT_cuda_synth_uniform": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64,
                                              C.c_uint16, C.c_void_p]),
    "FLAGSTAT_cuda_synth_hiseqx": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64,
                                             C.c_uint32, C.c_void_p]),
    "FLAGSTAT_cuda_malloc": (C.c_void_p, [C.c_size_t]),
    "FLAGSTAT_cuda_malloc_host": (C.c_void_p, [C.c_size_t]),
    "FLAGSTAT_cuda_free": (C.c_int, [C.c_void_p]),
    "FLAGSTAT_cuda_free_host": (C.c_int, [C.c_void_p]),
    "FLAGSTAT_cuda_memcpy_h2d": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "FLAGSTAT_cuda_memcpy_d2h": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "FLAGSTAT_cuda_memset": (C.c_int, [C.c_void_p, C.c_int, C.c_size_t]),
    "FLAGSTAT_cuda_sync": (C.c_int, []),
    "FLAGSTAT_cuda_time_device": (C.c_int, [C.c_void_p, C.c_uint64, C.c_void_p, C.c_int, C.c_int,
                                            C.POINTER(C.c_float)]),
    "FLAGSTAT_cuda_time_device_rot": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint32, C.c_void_p,
                                                C.c_int, C.c_int, C.POINTER(C.c_float)]),
    "FLAGSTAT_cuda_read_probe": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int, C.POINTER(C.c_float)]),
}

_lib = None


class FlagstatCudaError(RuntimeError):
    def __init__(self, code: int, what: str):
        self.code = code
        super().__init__(f"{what}: error {code}: {strerror(code)}")


def lib():
    """Load libflagstats_cuda.so (once).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise ImportError(
                f"{SO_PATH} is missing: build it with `python -m libflagstats_b200.build` "
                "(there is no CPU fallback)")
        handle = C.CDLL(SO_PATH, mode=C.RTLD_GLOBAL)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the ABI and the header drift apart
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def strerror(code: int) -> str:
    return lib().FLAGSTAT_cuda_strerror(int(code)).decode()


def check(code: int, what: str) -> None:
    if code != 0:
        raise FlagstatCudaError(code, what)
