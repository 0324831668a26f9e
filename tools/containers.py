"""Synthetic FLAG columns and the reference's block containers as bench / tool INPUT.

Writers of the files `bench compress` produces (benchmark/flagstats.cpp:110-215):
repeated [int32 raw_size][int32 comp_size][payload] records around 1,024,000-byte blocks,
payload = one LZ4 block (LZ4_compress_default / LZ4_compress_HC) or one Zstandard frame
(ZSTD_compress), made with the image's own liblz4.so.1 / libzstd.so.1 through ctypes (their
headers are not installed; these entry points have had the same signatures since lz4 1.7 /
zstd 1.0).  Nothing here comes from oracle/: this is input generation for the GPU legs.
"""
from __future__ import annotations

import ctypes as C
import os
import struct
from concurrent.futures import ThreadPoolExecutor

import numpy as np

BLOCK_BYTES = 1_024_000  # benchmark/flagstats.cpp:119

_lz4 = None
_zstd = None


def liblz4():
    global _lz4
    if _lz4 is None:
        try:
            z = C.CDLL("liblz4.so.1")
            z.LZ4_compressBound.restype = C.c_int
            z.LZ4_compressBound.argtypes = [C.c_int]
            z.LZ4_compress_default.restype = C.c_int
            z.LZ4_compress_default.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int]
            z.LZ4_compress_HC.restype = C.c_int
            z.LZ4_compress_HC.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int]
            _lz4 = z
        except (OSError, AttributeError):
            _lz4 = False
    return _lz4 or None


def libzstd():
    global _zstd
    if _zstd is None:
        try:
            z = C.CDLL("libzstd.so.1")
            z.ZSTD_compressBound.restype = C.c_size_t
            z.ZSTD_compressBound.argtypes = [C.c_size_t]
            z.ZSTD_compress.restype = C.c_size_t
            z.ZSTD_compress.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_int]
            z.ZSTD_isError.restype = C.c_uint
            z.ZSTD_isError.argtypes = [C.c_size_t]
            _zstd = z
        except (OSError, AttributeError):
            _zstd = False
    return _zstd or None


def lz4_block(raw: bytes, hc_level: int = 0) -> bytes:
    z = liblz4()
    cap = z.LZ4_compressBound(len(raw))
    dst = C.create_string_buffer(cap)
    n = (z.LZ4_compress_HC(raw, dst, len(raw), cap, hc_level) if hc_level
         else z.LZ4_compress_default(raw, dst, len(raw), cap))
    if n <= 0:
        raise RuntimeError("LZ4 compression failed")
    return dst.raw[:n]


def zstd_frame(raw: bytes, level: int = 1) -> bytes:
    z = libzstd()
    cap = z.ZSTD_compressBound(len(raw))
    dst = C.create_string_buffer(cap)
    n = z.ZSTD_compress(dst, cap, raw, len(raw), level)
    if z.ZSTD_isError(n):
        raise RuntimeError("ZSTD_compress failed")
    return dst.raw[:n]


def container(col: np.ndarray, codec: str, level: int = 0, threads: int = 0) -> bytes:
    """codec: 'lz4' (level 0 = LZ4_compress_default, else LZ4-HC level) or 'zstd'."""
    raw = np.ascontiguousarray(col, dtype=np.uint16).tobytes()
    chunks = [raw[lo:lo + BLOCK_BYTES] for lo in range(0, len(raw), BLOCK_BYTES)]
    fn = (lambda c: lz4_block(c, level)) if codec == "lz4" else (lambda c: zstd_frame(c, level or 1))
    threads = threads or len(os.sched_getaffinity(0))
    with ThreadPoolExecutor(threads) as ex:  # ctypes releases the GIL inside the codec
        comps = list(ex.map(fn, chunks))
    return b"".join(struct.pack("<ii", len(c), len(z)) + z for c, z in zip(chunks, comps))


HISEQX_CATS = np.array([99, 147, 83, 163, 97, 145, 73, 137, 133, 69, 77, 141, 2113, 2177], np.uint16)
HISEQX_P = np.array([195, 195, 195, 195, 8.4, 8.4, 1, 1, 1, 1, 8.5, 8.5, 1.3, 1.3])


def runs_column(n: int, seed: int = 1, mean_run: int = 8) -> np.ndarray:
    """HiSeqX FLAG categories (README.md:179-191 proportions) in geometric runs: coordinate-sorted
    files repeat flag patterns locally.  mean_run 8 compresses ~5x (LZ4) / ~12x (Zstd)."""
    rng = np.random.default_rng(seed)
    nruns = int(n / mean_run * 1.05) + 1000
    vals = rng.choice(HISEQX_CATS, size=nruns, p=HISEQX_P / HISEQX_P.sum())
    lens = rng.geometric(1.0 / mean_run, size=nruns)
    col = np.repeat(vals, lens)[:n].astype(np.uint16)
    assert col.size == n
    return col


def iid_column(n: int, seed: int = 3) -> np.ndarray:
    """The same categories drawn independently per record: LZ4 ratio ~2.2."""
    rng = np.random.default_rng(seed)
    return rng.choice(HISEQX_CATS, size=n, p=HISEQX_P / HISEQX_P.sum()).astype(np.uint16)
