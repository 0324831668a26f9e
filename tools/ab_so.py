#!/usr/bin/env python3
"""Same-box A/B of two builds of libflagstats_cuda.so (and of the plain vs fused launch).

    python tools/ab_so.py libflagstats_b200/libflagstats_cuda_prev.so libflagstats_b200/libflagstats_cuda.so
"""
import ctypes as C
import json
import sys

import torch

N = 824_541_892
ITERS = 400


def load(path):
    L = C.CDLL(path)
    L.FLAGSTAT_cuda_synth_hiseqx.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, C.c_void_p]
    L.FLAGSTAT_cuda_device.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]
    if hasattr(L, "FLAGSTAT_cuda_xchg_create"):
        L.FLAGSTAT_cuda_xchg_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_void_p]
        L.FLAGSTAT_cuda_device_allreduce.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_int, C.c_void_p]
    return L


def main():
    libs = [(p, load(p)) for p in sys.argv[1:]]
    data = torch.empty(N, dtype=torch.int16, device="cuda")
    libs[-1][1].FLAGSTAT_cuda_synth_hiseqx(data.data_ptr(), 0, N, 0, 0, None)
    torch.cuda.synchronize()
    out = torch.zeros(32, dtype=torch.int64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream

    def timed(fn):
        for _ in range(20):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(ITERS):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / ITERS * 1e3

    cases = []
    for p, L in libs:
        cases.append((p + " plain", lambda L=L: L.FLAGSTAT_cuda_device(data.data_ptr(), N, out.data_ptr(), st)))
        cases.append((p + " memset+plain", lambda L=L: (out.zero_(), L.FLAGSTAT_cuda_device(data.data_ptr(), N, out.data_ptr(), st))))
        if hasattr(L, "FLAGSTAT_cuda_xchg_create"):
            h = C.c_void_p()
            assert L.FLAGSTAT_cuda_xchg_create(C.byref(h), 0, 1, None) == 0
            cases.append((p + " fused(world=1)", lambda L=L, h=h: L.FLAGSTAT_cuda_device_allreduce(h, data.data_ptr(), N, out.data_ptr(), 0, st)))
    for rep in range(3):
        for name, fn in cases:
            print(json.dumps({"case": name, "rep": rep, "us_per_launch": round(timed(fn), 2)}), flush=True)


if __name__ == "__main__":
    main()
