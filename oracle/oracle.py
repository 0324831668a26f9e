"""ctypes front-end of the test oracle.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  Nothing under
libflagstats_b200/ may import this module.

Two back-ends:

* ``liboracle.so``  -- our C restatement (oracle/flagstat_oracle.c).  Built on
  demand with the host gcc; needs nothing outside this repo.
* ``oracle/_ref/libflagstats_ref_*.so`` -- the UNMODIFIED reference compiled
  from /root/reference by oracle/Makefile.  Prebuilt in the dev container and
  shipped to the GPU box; absent => ``reference()`` returns None and callers
  fall back to the restatement (kind "port").

A third, numpy-vectorised restatement (`numpy_flagstat`) is kept as an
independent cross-check that is fast enough for 10^8..10^9 records.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")

_u16p = C.POINTER(C.c_uint16)
_u32p = C.POINTER(C.c_uint32)
_u64p = C.POINTER(C.c_uint64)

CORE19 = (2, 6, 7, 8, 10, 11, 12, 13, 14, 18, 22, 23, 24, 25, 26, 27, 28, 29, 30)
"""Slots every correct reference kernel agrees on (SURVEY.md section 8a)."""
CORE20 = CORE19 + (9,)
"""CORE19 plus slot 9 under the SIMD convention -- what FLAGSTAT_cuda writes."""


def _ptr(a: np.ndarray, t):
    return a.ctypes.data_as(t)


def _as_u16(a) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint16)
    return a


# --------------------------------------------------------------------------
# restatement
# --------------------------------------------------------------------------
_oracle = None


def build_oracle(force: bool = False) -> str:
    so = os.path.join(HERE, "liboracle.so")
    srcs = [os.path.join(HERE, "flagstat_oracle.c"), os.path.join(HERE, "lz4_oracle.c"),
            os.path.join(HERE, "zstd_oracle.c")]
    if force or not os.path.exists(so) or any(os.path.getmtime(so) < os.path.getmtime(x) for x in srcs):
        subprocess.check_call(
            ["gcc", "-O2", "-std=c11", "-Wall", "-Wextra", "-fPIC", "-shared", "-o", so] + srcs
        )
    return so


def oracle():
    global _oracle
    if _oracle is None:
        lib = C.CDLL(build_oracle())
        for name in ("oracle_flagstat_scalar_u64", "oracle_flagstat_simd_u64",
                     "oracle_flagstat_maskselect_u64"):
            f = getattr(lib, name)
            f.argtypes = [_u16p, C.c_uint64, _u64p]
            f.restype = C.c_int
        for name in ("oracle_flagstat_scalar_u32", "oracle_flagstat_simd_u32"):
            f = getattr(lib, name)
            f.argtypes = [_u16p, C.c_uint32, _u32p]
            f.restype = C.c_int
        lib.oracle_samtools_loop.argtypes = [_u16p, C.c_uint64, C.POINTER(C.c_longlong)]
        lib.oracle_samtools_loop.restype = C.c_int
        lib.oracle_mask_select.argtypes = [C.c_uint16]
        lib.oracle_mask_select.restype = C.c_uint16
        lib.oracle_pospopcnt_u16_u64.argtypes = [_u16p, C.c_uint64, _u64p]
        lib.oracle_pospopcnt_u16.argtypes = [_u16p, C.c_size_t, _u32p]
        lib.oracle_synth_uniform.argtypes = [_u16p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint16]
        lib.oracle_synth_uniform.restype = None
        lib.oracle_synth_hiseqx.argtypes = [_u16p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32]
        lib.oracle_synth_hiseqx.restype = None
        lib.oracle_hiseqx_n.restype = C.c_uint64
        lib.oracle_hiseqx_m.restype = C.c_uint64
        lib.oracle_lz4_decompress.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64]
        lib.oracle_lz4_decompress.restype = C.c_int64
        lib.oracle_lz4_compress.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64]
        lib.oracle_lz4_compress.restype = C.c_int64
        lib.oracle_zstd_decompress.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64]
        lib.oracle_zstd_decompress.restype = C.c_int64
        _oracle = lib
    return _oracle


def flagstat_scalar(a, flags: Optional[np.ndarray] = None) -> np.ndarray:
    """Scalar convention, libflagstats.h:170-176 (slot 9 untouched)."""
    a = _as_u16(a)
    f = np.zeros(32, np.uint64) if flags is None else flags
    oracle().oracle_flagstat_scalar_u64(_ptr(a, _u16p), a.size, _ptr(f, _u64p))
    return f


def flagstat_simd(a, flags: Optional[np.ndarray] = None) -> np.ndarray:
    """SIMD convention (slot 9 += n_pass), libflagstats.h:429,1212,1843."""
    a = _as_u16(a)
    f = np.zeros(32, np.uint64) if flags is None else flags
    oracle().oracle_flagstat_simd_u64(_ptr(a, _u16p), a.size, _ptr(f, _u64p))
    return f


def flagstat_simd_u32(a, flags: Optional[np.ndarray] = None) -> np.ndarray:
    a = _as_u16(a)
    f = np.zeros(32, np.uint32) if flags is None else flags
    oracle().oracle_flagstat_simd_u32(_ptr(a, _u16p), a.size, _ptr(f, _u32p))
    return f


def flagstat_maskselect(a) -> np.ndarray:
    a = _as_u16(a)
    f = np.zeros(32, np.uint64)
    oracle().oracle_flagstat_maskselect_u64(_ptr(a, _u16p), a.size, _ptr(f, _u64p))
    return f


def mask_select(v: int) -> int:
    return int(oracle().oracle_mask_select(int(v) & 0xFFFF))


SAMTOOLS_FIELDS = ("n_reads", "n_mapped", "n_pair_all", "n_pair_map", "n_pair_good", "n_sgltn",
                   "n_read1", "n_read2", "n_dup", "n_diffchr", "n_diffhigh", "n_secondary", "n_supp")
"""bam_flagstat_t in declaration order, benchmark/flagstats.cpp:43-49; each field is [pass, fail]."""


def samtools_loop(a, s: Optional[np.ndarray] = None) -> np.ndarray:
    """flagstat_loop (benchmark/flagstats.cpp:51-71) over a column; int64[13, 2], accumulates."""
    a = _as_u16(a)
    s = np.zeros((13, 2), np.int64) if s is None else s
    oracle().oracle_samtools_loop(_ptr(a, _u16p), a.size, _ptr(s, C.POINTER(C.c_longlong)))
    return s


def samtools_from_counters(flags) -> np.ndarray:
    """bam_flagstat_t implied by the 32 flagstat counters (SIMD convention) with the exact
    n_pair_all in slots 0 / 16 -- the identities FLAGSTAT_cuda_samtools relies on."""
    f = [int(x) for x in np.asarray(flags)]
    s = np.zeros((13, 2), np.int64)
    for w in (0, 1):
        o = 16 * w
        total = f[9] if w == 0 else f[25]
        s[0, w] = total
        s[1, w] = total - f[o + 2]
        s[2, w] = f[o + 0]
        s[3, w] = f[o + 14]
        s[4, w] = f[o + 12]
        s[5, w] = f[o + 13]
        s[6, w] = f[o + 6]
        s[7, w] = f[o + 7]
        s[8, w] = f[o + 10]
        s[11, w] = f[o + 8]
        s[12, w] = f[o + 11]
    return s


def samtools_percent(n: int, total: int) -> str:
    """percent(), benchmark/flagstats.cpp:73-78: "%.2f%%" of (float)n / total * 100.0."""
    if total == 0:
        return "N/A"
    return "%.2f%%" % (float(np.float32(n) / np.float32(total)) * 100.0)


def samtools_report(s) -> str:
    """The report of benchmark/flagstats.cpp:577-588 (diffchr lines are commented out there)."""
    s = np.asarray(s).reshape(13, 2)
    g = {k: (int(s[i, 0]), int(s[i, 1])) for i, k in enumerate(SAMTOOLS_FIELDS)}
    pc = samtools_percent
    lines = [
        "%d + %d in total (QC-passed reads + QC-failed reads)" % g["n_reads"],
        "%d + %d secondary" % g["n_secondary"],
        "%d + %d supplementary" % g["n_supp"],
        "%d + %d duplicates" % g["n_dup"],
        "%d + %d mapped (%s : %s)" % (g["n_mapped"] + (pc(g["n_mapped"][0], g["n_reads"][0]),
                                                        pc(g["n_mapped"][1], g["n_reads"][1]))),
        "%d + %d paired in sequencing" % g["n_pair_all"],
        "%d + %d read1" % g["n_read1"],
        "%d + %d read2" % g["n_read2"],
        "%d + %d properly paired (%s : %s)" % (g["n_pair_good"] + (
            pc(g["n_pair_good"][0], g["n_pair_all"][0]), pc(g["n_pair_good"][1], g["n_pair_all"][1]))),
        "%d + %d with itself and mate mapped" % g["n_pair_map"],
        "%d + %d singletons (%s : %s)" % (g["n_sgltn"] + (
            pc(g["n_sgltn"][0], g["n_pair_all"][0]), pc(g["n_sgltn"][1], g["n_pair_all"][1]))),
    ]
    return "\n".join(lines) + "\n"


def pospopcnt(a) -> np.ndarray:
    a = _as_u16(a)
    out = np.empty(16, np.uint64)
    oracle().oracle_pospopcnt_u16_u64(_ptr(a, _u16p), a.size, _ptr(out, _u64p))
    return out


def synth_uniform(start: int, n: int, seed: int = 0, mask: int = 0x0FFF) -> np.ndarray:
    out = np.empty(n, np.uint16)
    oracle().oracle_synth_uniform(_ptr(out, _u16p), start, n, seed, mask)
    return out


def synth_hiseqx(start: int, n: int, seed: int = 0, qcfail_ppm: int = 0) -> np.ndarray:
    out = np.empty(n, np.uint16)
    oracle().oracle_synth_hiseqx(_ptr(out, _u16p), start, n, seed, qcfail_ppm)
    return out


HISEQX_N = 824_541_892


def numpy_flagstat(a) -> np.ndarray:
    """Vectorised restatement of libflagstats.h:118-142 + the slot-9 rule
    (:429).  Independent of the C code above; used as a cross-check and for
    arrays too large for the scalar loop."""
    a = _as_u16(a)
    out = np.zeros(32, np.uint64)
    step = 1 << 24
    for lo in range(0, a.size, step):
        x = a[lo:lo + step]
        bit = lambda b: (x >> b) & 1 == 1  # noqa: E731
        fail = bit(9)
        sec, supp, paired = bit(8), bit(11), bit(0)
        unmap, munmap = bit(2), bit(3)
        third = ~sec & ~supp & paired
        ind = {
            2: unmap, 10: bit(10), 8: sec, 11: ~sec & supp,
            12: third & bit(1) & ~unmap, 6: third & bit(6), 7: third & bit(7),
            13: third & munmap & ~unmap, 14: third & ~unmap & ~munmap,
        }
        for j, m in ind.items():
            nf = int(np.count_nonzero(m & fail))
            out[16 + j] += np.uint64(nf)
            out[j] += np.uint64(int(np.count_nonzero(m)) - nf)
        nfail = int(np.count_nonzero(fail))
        out[25] += np.uint64(nfail)
        out[9] += np.uint64(x.size - nfail)
    return out


# --------------------------------------------------------------------------
# the real reference, when its prebuilt shim is available
# --------------------------------------------------------------------------
_ref = None
_ref_tried = False


def _cpu_flags() -> set:
    try:
        with open("/proc/cpuinfo") as fh:
            for line in fh:
                if line.startswith("flags"):
                    return set(line.split(":", 1)[1].split())
    except OSError:
        pass
    return set()


_V3 = {"avx", "avx2", "bmi1", "bmi2", "f16c", "fma", "abm", "movbe"}
_V4 = _V3 | {"avx512f", "avx512bw", "avx512cd", "avx512dq", "avx512vl"}


def reference_path() -> Optional[str]:
    flags = _cpu_flags()
    order = []
    if _V4 <= flags:
        order.append("v4")
    if _V3 <= flags:
        order.append("v3")
    order.append("generic")
    for v in order:
        p = os.path.join(REF_DIR, f"libflagstats_ref_{v}.so")
        if os.path.exists(p):
            return p
    return None


def build_reference() -> bool:
    """(Re)build oracle/_ref from /root/reference when that tree is present."""
    if not os.path.exists("/root/reference/libflagstats.h"):
        return False
    subprocess.check_call(["make", "-s", "-C", HERE, "ref"])
    return True


def reference():
    """ctypes handle of the unmodified reference, or None if not prebuilt."""
    global _ref, _ref_tried
    if _ref_tried:
        return _ref
    _ref_tried = True
    p = reference_path()
    if p is None:
        return None
    lib = C.CDLL(p)
    lib.ref_flagstat.argtypes = [C.c_char_p, _u16p, C.c_uint32, _u32p]
    lib.ref_flagstat.restype = C.c_int
    lib.ref_FLAGSTATS_u16.argtypes = [_u16p, C.c_uint32, _u32p]
    lib.ref_FLAGSTATS_u16.restype = C.c_uint64
    lib.ref_dispatch_name.argtypes = [C.c_uint32]
    lib.ref_dispatch_name.restype = C.c_char_p
    lib.ref_pospopcnt_u16.argtypes = [_u16p, C.c_size_t, _u32p]
    lib.ref_pospopcnt_u16.restype = C.c_int
    lib.ref_kernel_name.argtypes = [C.c_int]
    lib.ref_kernel_name.restype = C.c_char_p
    lib.ref_kernel_runnable.argtypes = [C.c_char_p]
    lib.ref_flagstat_mt.argtypes = [C.c_char_p, _u16p, C.c_uint64, C.c_int, _u64p,
                                    C.POINTER(C.c_double)]
    lib.ref_flagstat_mt.restype = C.c_int
    if hasattr(lib, "ref_pospopcnt_mt"):
        lib.ref_pospopcnt_mt.argtypes = [_u16p, C.c_uint64, C.c_int, _u64p, C.POINTER(C.c_double)]
        lib.ref_pospopcnt_mt.restype = C.c_int
        lib.ref_codec_available.argtypes = [C.c_int]
        lib.ref_codec_available.restype = C.c_int
        lib.ref_container_mt.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_int, _u64p, _u64p,
                                         C.POINTER(C.c_double), C.POINTER(C.c_double)]
        lib.ref_container_mt.restype = C.c_int
    if hasattr(lib, "ref_samtools_loop"):
        lib.ref_samtools_loop.argtypes = [_u16p, C.c_uint64, C.POINTER(C.c_longlong)]
        lib.ref_samtools_loop.restype = C.c_int
        lib.ref_samtools_percent.argtypes = [C.c_longlong, C.c_longlong, C.c_char_p]
        lib.ref_samtools_percent.restype = C.c_int
    lib._path = p
    _ref = lib
    return _ref


REF_CORRECT_KERNELS = ("scalar", "sse4", "sse4_improved", "sse4_improved2", "avx2", "avx512",
                       "avx512_improved", "avx512_improved2", "avx512_improved3")
"""Kernels that agree with FLAGSTAT_scalar on CORE19 (SURVEY.md section 2.2);
avx2_improved{,2} and avx512_improved4 are reference defects."""


def ref_kernels(runnable_only: bool = True):
    lib = reference()
    if lib is None:
        return []
    names = [lib.ref_kernel_name(i).decode() for i in range(lib.ref_num_kernels())]
    if runnable_only:
        names = [n for n in names if lib.ref_kernel_runnable(n.encode())]
    return names


def ref_flagstat(name: str, a, flags: Optional[np.ndarray] = None) -> np.ndarray:
    lib = reference()
    a = _as_u16(a)
    f = np.zeros(32, np.uint32) if flags is None else flags
    rc = lib.ref_flagstat(name.encode(), _ptr(a, _u16p), a.size, _ptr(f, _u32p))
    if rc != 0:
        raise RuntimeError(f"reference kernel {name!r}: rc={rc}")
    return f


def ref_flagstats_u16(a, flags: Optional[np.ndarray] = None) -> np.ndarray:
    lib = reference()
    a = _as_u16(a)
    f = np.zeros(32, np.uint32) if flags is None else flags
    lib.ref_FLAGSTATS_u16(_ptr(a, _u16p), a.size, _ptr(f, _u32p))
    return f


def ref_samtools_loop(a) -> Optional[np.ndarray]:
    """The reference's own flagstat_loop macro (benchmark/flagstats.cpp:51-71) compiled into
    oracle/_ref; None when the prebuilt shim predates it."""
    lib = reference()
    if lib is None or not hasattr(lib, "ref_samtools_loop"):
        return None
    a = _as_u16(a)
    s = np.zeros((13, 2), np.int64)
    if lib.ref_samtools_loop(_ptr(a, _u16p), a.size, _ptr(s, C.POINTER(C.c_longlong))) != 0:
        return None
    return s


def ref_samtools_percent(n: int, total: int) -> Optional[str]:
    lib = reference()
    if lib is None or not hasattr(lib, "ref_samtools_percent"):
        return None
    buf = C.create_string_buffer(64)
    if lib.ref_samtools_percent(n, total, buf) != 0:
        return None
    return buf.value.decode()


def ref_dispatch_name(n: int) -> str:
    return reference().ref_dispatch_name(n).decode()


def ref_pospopcnt(a) -> np.ndarray:
    lib = reference()
    a = _as_u16(a)
    out = np.full(16, 0xDEADBEEF, np.uint32)  # the reference memsets it
    lib.ref_pospopcnt_u16(_ptr(a, _u16p), a.size, _ptr(out, _u32p))
    return out


def ref_flagstat_mt(name: str, a, nthreads: int):
    """(flags u64[32], seconds) of the range-sharded pthread wrapper."""
    lib = reference()
    a = _as_u16(a)
    f = np.zeros(32, np.uint64)
    sec = C.c_double(0.0)
    rc = lib.ref_flagstat_mt(name.encode(), _ptr(a, _u16p), a.size, nthreads,
                             _ptr(f, _u64p), C.byref(sec))
    if rc != 0:
        raise RuntimeError(f"reference kernel {name!r}: rc={rc}")
    return f, sec.value


def ref_pospopcnt_mt(a, nthreads: int):
    """(out u64[16], seconds): STORM_pospopcnt_u16 over contiguous ranges on pthreads."""
    lib = reference()
    a = _as_u16(a)
    out = np.zeros(16, np.uint64)
    sec = C.c_double(0.0)
    rc = lib.ref_pospopcnt_mt(_ptr(a, _u16p), a.size, nthreads, _ptr(out, _u64p), C.byref(sec))
    if rc != 0:
        raise RuntimeError(f"ref_pospopcnt_mt: rc={rc}")
    return out, sec.value


REF_CODECS = {"lz4": 0, "zstd": 1}


def ref_container_available(codec: str) -> bool:
    lib = reference()
    return bool(lib is not None and hasattr(lib, "ref_container_mt")
                and lib.ref_codec_available(REF_CODECS[codec]))


def ref_container_mt(blob, codec: str, nthreads: int):
    """The reference's block loop (benchmark/flagstats.cpp:311-331 LZ4, 655-669 Zstd:
    system codec, then FLAGSTATS_get_function(N) per block) over a container in memory.
    Returns (flags u64[32], n_records, seconds, decode_seconds)."""
    lib = reference()
    buf = np.frombuffer(blob, dtype=np.uint8)
    f = np.zeros(32, np.uint64)
    n = C.c_uint64(0)
    sec, dec = C.c_double(0.0), C.c_double(0.0)
    rc = lib.ref_container_mt(buf.ctypes.data, buf.size, REF_CODECS[codec], nthreads, _ptr(f, _u64p),
                              C.byref(n), C.byref(sec), C.byref(dec))
    if rc != 0:
        raise RuntimeError(f"ref_container_mt({codec}): rc={rc}")
    return f, int(n.value), sec.value, dec.value


def best_reference_kernel() -> Optional[str]:
    """What FLAGSTATS_get_function returns for a large block on this CPU."""
    if reference() is None:
        return None
    return ref_dispatch_name(1 << 20)


# --------------------------------------------------------------------------
# LZ4 block containers (the caller either side of the path)
# --------------------------------------------------------------------------
REF_BLOCK_BYTES = 1_024_000
"""benchmark/flagstats.cpp:119 -- 512,000 records per block."""


def lz4_decompress(block: bytes, raw_size: int) -> bytes:
    """oracle/lz4_oracle.c restatement of LZ4_decompress_safe; raises on malformed input."""
    out = C.create_string_buffer(max(raw_size, 1))
    got = oracle().oracle_lz4_decompress(block, len(block), out, raw_size)
    if got != raw_size:
        raise ValueError(f"malformed LZ4 block (decoder returned {got}, expected {raw_size})")
    return out.raw[:raw_size]


def lz4_compress(raw: bytes) -> bytes:
    """Greedy encoder of oracle/lz4_oracle.c (valid LZ4 blocks, not liblz4's choices)."""
    cap = len(raw) + len(raw) // 255 + 64
    out = C.create_string_buffer(cap)
    got = oracle().oracle_lz4_compress(raw, len(raw), out, cap)
    if got < 0:
        raise ValueError("compress bound exceeded")
    return out.raw[:got]


def liblz4_compress(raw: bytes) -> bytes:
    """A real liblz4 (through pyarrow's lz4_raw codec): what the reference's lz4f() calls."""
    import pyarrow as pa

    return pa.compress(raw, codec="lz4_raw", asbytes=True)


def liblz4_decompress(block: bytes, raw_size: int) -> bytes:
    import pyarrow as pa

    return pa.decompress(block, decompressed_size=raw_size, codec="lz4_raw", asbytes=True)


def write_lz4_container(a, block_bytes: int = REF_BLOCK_BYTES, compressor=None) -> bytes:
    """The file lz4f() writes (benchmark/flagstats.cpp:110-147) for the FLAG column `a`."""
    import struct

    compressor = compressor or liblz4_compress
    raw = _as_u16(a).tobytes()
    out = []
    for lo in range(0, len(raw), block_bytes):
        chunk = raw[lo:lo + block_bytes]
        comp = compressor(chunk)
        out.append(struct.pack("<ii", len(chunk), len(comp)))
        out.append(comp)
    return b"".join(out)


def read_lz4_container(blob: bytes, decompressor=None):
    """Block loop of benchmark/flagstats.cpp:288-358: yields uint16 arrays, one per block."""
    import struct

    decompressor = decompressor or lz4_decompress
    pos = 0
    while pos < len(blob):
        raw_size, comp_size = struct.unpack_from("<ii", blob, pos)
        pos += 8
        chunk = decompressor(blob[pos:pos + comp_size], raw_size)
        pos += comp_size
        yield np.frombuffer(chunk[: (raw_size >> 1) * 2], dtype=np.uint16)


# --------------------------------------------------------------------------
# Zstd block containers (benchmark/flagstats.cpp:192-215 writer, :636-676 reader)
# --------------------------------------------------------------------------
def zstd_decompress(frame: bytes, raw_size: int) -> bytes:
    """oracle/zstd_oracle.c restatement of ZSTD_decompress for one frame (RFC 8878)."""
    out = C.create_string_buffer(max(raw_size, 1))
    got = oracle().oracle_zstd_decompress(frame, len(frame), out, raw_size)
    if got != raw_size:
        raise ValueError(f"malformed Zstd frame (decoder returned {got}, expected {raw_size})")
    return out.raw[:raw_size]


_libzstd = None


def libzstd():
    """The real libzstd runtime of this image (libzstd.so.1; no headers are installed, the three
    entry points used here have had this signature since zstd 1.0), or None."""
    global _libzstd
    if _libzstd is None:
        try:
            z = C.CDLL("libzstd.so.1")
        except OSError:
            _libzstd = False
            return None
        z.ZSTD_compressBound.restype = C.c_size_t
        z.ZSTD_compressBound.argtypes = [C.c_size_t]
        z.ZSTD_compress.restype = C.c_size_t
        z.ZSTD_compress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int]
        z.ZSTD_decompress.restype = C.c_size_t
        z.ZSTD_decompress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        z.ZSTD_isError.restype = C.c_uint
        z.ZSTD_isError.argtypes = [C.c_size_t]
        z.ZSTD_versionString.restype = C.c_char_p
        _libzstd = z
    return _libzstd or None


def libzstd_compress(raw: bytes, level: int = 1) -> bytes:
    """ZSTD_compress(out, cap, in, n, level): what the reference's ZstdCompress calls (:90-93)."""
    z = libzstd()
    cap = z.ZSTD_compressBound(len(raw))
    dst = C.create_string_buffer(max(cap, 1))
    n = z.ZSTD_compress(dst, cap, raw, len(raw), level)
    if z.ZSTD_isError(n):
        raise ValueError("ZSTD_compress failed")
    return dst.raw[:n]


def libzstd_decompress(frame: bytes, raw_size: int) -> bytes:
    z = libzstd()
    dst = C.create_string_buffer(max(raw_size, 1))
    n = z.ZSTD_decompress(dst, raw_size, frame, len(frame))
    if z.ZSTD_isError(n) or n != raw_size:
        raise ValueError("ZSTD_decompress failed")
    return dst.raw[:raw_size]


def write_zstd_container(a, level: int = 1, block_bytes: int = REF_BLOCK_BYTES) -> bytes:
    """The file zstd() writes (benchmark/flagstats.cpp:192-215) for the FLAG column `a`."""
    import struct

    raw = _as_u16(a).tobytes()
    out = []
    for lo in range(0, len(raw), block_bytes):
        chunk = raw[lo:lo + block_bytes]
        comp = libzstd_compress(chunk, level)
        out.append(struct.pack("<ii", len(chunk), len(comp)))
        out.append(comp)
    return b"".join(out)


def read_zstd_container(blob: bytes, decompressor=None):
    """Block loop of zstd_decompress() (benchmark/flagstats.cpp:636-676)."""
    return read_lz4_container(blob, decompressor or zstd_decompress)


# --------------------------------------------------------------------------
# FLAG ingest (benchmark/utility.cpp:29-32)
# --------------------------------------------------------------------------
def ingest_text(text: bytes) -> np.ndarray:
    """while (getline(cin, str)) write((uint16_t) atoi(str.c_str())) -- with the C library's
    own atoi, one call per line."""
    libc = C.CDLL(None)
    libc.atoi.argtypes = [C.c_char_p]
    libc.atoi.restype = C.c_int
    lines = text.split(b"\n")
    if lines and lines[-1] == b"":
        lines.pop()  # getline: nothing after the final newline is not a line
    return np.array([libc.atoi(ln) & 0xFFFF for ln in lines], dtype=np.uint16)
