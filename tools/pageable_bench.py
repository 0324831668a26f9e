#!/usr/bin/env python3
"""FLAGSTAT_cuda_u64 on PAGEABLE host memory (what numpy / malloc give the reference's callers,
python/libflagstats.pyx:22): staging-thread sweep, and in-place cudaHostRegister as the
alternative for large arrays.  JSON lines.   python tools/pageable_bench.py [records]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import libflagstats_b200 as fs  # noqa: E402
from libflagstats_b200 import synth  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else synth.HISEQX_N
    d = synth.hiseqx_device(n)
    page = np.empty(n, np.uint16)
    page[:] = d.cpu().numpy().view(np.uint16)
    pin = torch.empty(n, dtype=torch.int16, pin_memory=True)
    pin.copy_(d)
    pin_np = pin.numpy().view(np.uint16)
    want = fs.flagstat_u64(pin_np)
    t0 = time.perf_counter()
    for _ in range(3):
        fs.flagstat_u64(pin_np)
    pinned_s = (time.perf_counter() - t0) / 3
    print(json.dumps({"case": "pinned", "records": n, "gbs": 2 * n / pinned_s / 1e9, "cpus": len(os.sched_getaffinity(0))}), flush=True)
    for T in (1, 2, 4, 6, 8, 10, 12, 14, 16, 20, 24, 28, 32):
        if T > len(os.sched_getaffinity(0)) or T > 24:
            break
        os.environ["FLAGSTAT_CUDA_IO_THREADS"] = str(T)
        f = fs.flagstat_u64(page)
        best = 1e30
        for _ in range(3):
            t0 = time.perf_counter()
            f = fs.flagstat_u64(page)
            best = min(best, time.perf_counter() - t0)
        print(json.dumps({"case": "pageable_staged", "threads": T, "gbs": 2 * n / best / 1e9,
                          "frac_of_pinned": pinned_s / best, "same": f.tolist() == want.tolist()}), flush=True)
    os.environ.pop("FLAGSTAT_CUDA_IO_THREADS", None)
    # in-place registration of the caller's pages
    rt = torch.cuda.cudart()
    for rep in range(3):
        t0 = time.perf_counter()
        rc = rt.cudaHostRegister(page.ctypes.data, page.nbytes, 0)
        t1 = time.perf_counter()
        f = fs.flagstat_u64(page)
        t2 = time.perf_counter()
        rt.cudaHostUnregister(page.ctypes.data)
        t3 = time.perf_counter()
        print(json.dumps({"case": "host_register_in_place", "rc": int(rc), "register_s": t1 - t0, "call_s": t2 - t1,
                          "unregister_s": t3 - t2, "gbs_incl_register": 2 * n / (t3 - t0) / 1e9,
                          "gbs_call_only": 2 * n / (t2 - t1) / 1e9, "same": f.tolist() == want.tolist()}), flush=True)


if __name__ == "__main__":
    main()
