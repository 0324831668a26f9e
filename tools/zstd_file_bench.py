#!/usr/bin/env python3
"""End-to-end timing of the reference's Zstd block containers through the GPU
(benchmark/flagstats.cpp:192-215 writer, :636-676 reader): frames written by the real libzstd,
page-cache resident file, decoded on the GPU and counted from HBM; beside it the shape of the
reference's own loop on one host core (libzstd decode + the oracle's counter, first 100 blocks).

    python tools/zstd_file_bench.py [n_blocks] [levels...] > gpurun_out/zstd_file_bench.jsonl
"""
import json
import os
import struct
import sys
import tempfile
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import libflagstats_b200 as fs  # noqa: E402
from libflagstats_b200 import blockfile  # noqa: E402
from oracle import oracle as O  # noqa: E402  (frame writer, checker and the CPU leg only)
from file_bench import runs_column  # noqa: E402


def write_container(col, level, threads):
    raw = col.tobytes()
    chunks = [raw[lo:lo + O.REF_BLOCK_BYTES] for lo in range(0, len(raw), O.REF_BLOCK_BYTES)]
    with ThreadPoolExecutor(threads) as ex:  # ctypes releases the GIL inside ZSTD_compress
        comps = list(ex.map(lambda c: O.libzstd_compress(c, level), chunks))
    return b"".join(struct.pack("<ii", len(c), len(z)) + z for c, z in zip(chunks, comps))


def main():
    n_blocks = int(sys.argv[1]) if len(sys.argv) > 1 else 1600
    levels = [int(x) for x in sys.argv[2:]] or [1, 3]
    n = n_blocks * fs.BLOCK_RECORDS + 12_345
    threads = len(os.sched_getaffinity(0))
    tmp = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    col = runs_column(n)
    want = O.numpy_flagstat(col).tolist()
    for level in levels:
        t0 = time.perf_counter()
        blob = write_container(col, level, threads)
        t_comp = time.perf_counter() - t0
        path = os.path.join(tmp, f"flags_l{level}.zst")
        with open(path, "wb") as fh:
            fh.write(blob)
        best = None
        for _ in range(3):
            t0 = time.perf_counter()
            f, got_n = blockfile.flagstat_file(path)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        assert got_n == n and f.tolist() == want, level
        print(json.dumps({"column": "hiseqx_categories_runs_mean8", "file": f"zstd container, level {level}, GPU decode "
                          "(one thread per frame)", "records": n, "file_bytes": len(blob),
                          "ratio": round(2 * n / len(blob), 2), "best_s": round(best, 4),
                          "grec_s": round(n / best / 1e9, 2), "record_gbs": round(2 * n / best / 1e9, 2),
                          "verified": True, "libzstd_compress_s": round(t_comp, 2), "compress_threads": threads}),
              flush=True)
        t0 = time.perf_counter()
        fl = np.zeros(32, np.uint64)
        nrec = pos = 0
        for _ in range(min(100, n_blocks)):
            raw_size, comp_size = struct.unpack_from("<ii", blob, pos)
            pos += 8
            b = np.frombuffer(O.libzstd_decompress(blob[pos:pos + comp_size], raw_size), dtype=np.uint16)
            pos += comp_size
            O.flagstat_simd(b, fl)
            nrec += b.size
        dt = time.perf_counter() - t0
        print(json.dumps({"column": "hiseqx_categories_runs_mean8", "file": f"zstd container, level {level}, CPU: libzstd "
                          "decode + oracle counter, 1 thread, first 100 blocks", "records": nrec,
                          "best_s": round(dt, 4), "grec_s": round(nrec / dt / 1e9, 3)}), flush=True)
        os.remove(path)
    os.rmdir(tmp)


if __name__ == "__main__":
    main()
