#!/usr/bin/env python3
"""End-to-end timing of the reference's FLAG files through the GPU:
raw .bin and LZ4 block containers (liblz4-compressed via pyarrow), page-cache resident.

    python tools/file_bench.py [n_blocks] > gpurun_out/file_bench.jsonl
"""
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import libflagstats_b200 as fs  # noqa: E402
from libflagstats_b200 import blockfile, synth  # noqa: E402
from oracle import oracle as O  # noqa: E402  (checker + the CPU leg only)


def runs_column(n, seed=1, mean_run=8):
    """HiSeqX categories in runs (coordinate-sorted files repeat flag patterns locally)."""
    rng = np.random.default_rng(seed)
    cats = np.array([99, 147, 83, 163, 97, 145, 73, 137, 133, 69, 77, 141, 2113, 2177], np.uint16)
    p = np.array([195, 195, 195, 195, 8.4, 8.4, 1, 1, 1, 1, 8.5, 8.5, 1.3, 1.3])
    nruns = int(n / mean_run * 1.05) + 1000
    vals = rng.choice(cats, size=nruns, p=p / p.sum())
    lens = rng.geometric(1.0 / mean_run, size=nruns)
    col = np.repeat(vals, lens)[:n].astype(np.uint16)
    assert col.size == n
    return col


def main():
    n_blocks = int(sys.argv[1]) if len(sys.argv) > 1 else 1600
    n = n_blocks * fs.BLOCK_RECORDS + 12_345
    tmp = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    rng = np.random.default_rng(3)
    cats = np.array([99, 147, 83, 163, 97, 145, 73, 137, 133, 69, 77, 141, 2113, 2177], np.uint16)
    p = np.array([195, 195, 195, 195, 8.4, 8.4, 1, 1, 1, 1, 8.5, 8.5, 1.3, 1.3])
    cols = {
        "hiseqx_generator (quasi-periodic)": synth.hiseqx_device(n, 0, 0, 0).cpu().numpy().view(np.uint16),
        "hiseqx_categories_runs_mean8": runs_column(n),
        "hiseqx_categories_iid": rng.choice(cats, size=n, p=p / p.sum()).astype(np.uint16),
    }
    for name, col in cols.items():
        want = O.numpy_flagstat(col).tolist()
        p_raw = os.path.join(tmp, name + ".bin")
        col.tofile(p_raw)
        t0 = time.perf_counter()
        blob = O.write_lz4_container(col)
        t_comp = time.perf_counter() - t0
        p_lz4 = os.path.join(tmp, name + ".lz4")
        with open(p_lz4, "wb") as fh:
            fh.write(blob)
        for label, path, variant in (("raw .bin", p_raw, None),
                                     ("lz4 container, group decoder (32 sequences per warp step)", p_lz4, 1),
                                     ("lz4 container, sequence decoder (1 sequence per warp step)", p_lz4, 0)):
            if variant is not None:
                fs.lib().FLAGSTAT_cuda_set_lz4_variant(variant)
            best = None
            for _ in range(4):
                t0 = time.perf_counter()
                f, got_n = blockfile.flagstat_file(path)
                dt = time.perf_counter() - t0
                best = dt if best is None else min(best, dt)
            assert got_n == n and f.tolist() == want, (name, label)
            print(json.dumps({"column": name, "file": label, "records": n, "file_bytes": os.path.getsize(path),
                              "ratio": round(2 * n / os.path.getsize(path), 2), "best_s": round(best, 4),
                              "grec_s": round(n / best / 1e9, 2), "record_gbs": round(2 * n / best / 1e9, 2),
                              "verified": True}), flush=True)
        # CPU leg on a bounded sample (first 100 blocks): a real liblz4 (pyarrow) decoding each block,
        # then the oracle's counter -- the shape of the reference's loop, flagstats.cpp:288-358, 1 thread
        import struct
        t0 = time.perf_counter()
        f = np.zeros(32, np.uint64)
        nrec = pos = 0
        for _ in range(min(100, n_blocks)):
            raw_size, comp_size = struct.unpack_from("<ii", blob, pos)
            pos += 8
            b = np.frombuffer(O.liblz4_decompress(blob[pos:pos + comp_size], raw_size), dtype=np.uint16)
            pos += comp_size
            O.flagstat_simd(b, f)
            nrec += b.size
        dt = time.perf_counter() - t0
        print(json.dumps({"column": name, "file": "lz4 container, CPU: liblz4 decode + oracle counter, 1 thread, first 100 blocks",
                          "records": nrec, "best_s": round(dt, 4), "grec_s": round(nrec / dt / 1e9, 3),
                          "liblz4_compress_s_whole_column": round(t_comp, 2)}), flush=True)
        os.remove(p_raw)
        os.remove(p_lz4)
    os.rmdir(tmp)


if __name__ == "__main__":
    main()
