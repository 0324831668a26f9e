#!/usr/bin/env python3
"""Raw .bin reader (FLAGSTAT_cuda_file_u64, page-cache resident file): slot size x reader threads, at
the bench's 410 MB file and at the 1.65 GB column.  JSON lines.   python tools/raw_reader_sweep.py"""
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import libflagstats_b200 as fs  # noqa: E402
from libflagstats_b200 import blockfile, synth  # noqa: E402


def main():
    tmp = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    host = torch.empty(synth.HISEQX_N, dtype=torch.int16, pin_memory=True)
    host.copy_(synth.hiseqx_device(synth.HISEQX_N))
    torch.cuda.synchronize()
    full = host.numpy().view(np.uint16)
    t0 = time.perf_counter()
    for _ in range(3):
        fs.flagstat_u64(full)
    pinned = 2 * full.size / ((time.perf_counter() - t0) / 3) / 1e9
    print(json.dumps({"case": "pinned array through FLAGSTAT_cuda_u64", "gbs": pinned}), flush=True)
    cpus = len(os.sched_getaffinity(0))
    for n in (204_812_345, synth.HISEQX_N):
        path = os.path.join(tmp, f"flags_{n}.bin")
        full[:n].tofile(path)
        want = None
        for threads in sorted({4, 8, 12, min(16, cpus), min(24, cpus)}):
            if threads > cpus:
                continue
            for slot_kb in (0, 512, 1024, 2048, 4096, 8000):
                os.environ["FLAGSTAT_CUDA_IO_THREADS"] = str(threads)
                if slot_kb:
                    os.environ["FLAGSTAT_CUDA_RAW_SLOT_KB"] = str(slot_kb)
                else:
                    os.environ.pop("FLAGSTAT_CUDA_RAW_SLOT_KB", None)
                best = 1e30
                for _ in range(5):
                    t0 = time.perf_counter()
                    f, got = blockfile.flagstat_file(path)
                    best = min(best, time.perf_counter() - t0)
                want = want if want is not None else f.tolist()
                print(json.dumps({"case": "raw .bin", "records": n, "threads": threads, "slot_kb": slot_kb or "default",
                                  "ms": round(best * 1e3, 3), "gbs": round(2 * n / best / 1e9, 2),
                                  "frac_of_pinned": round(2 * n / best / 1e9 / pinned, 3),
                                  "same": f.tolist() == want and got == n}), flush=True)
        os.remove(path)
    os.rmdir(tmp)


if __name__ == "__main__":
    main()
