#!/usr/bin/env python3
"""Where does one launch of the flagstat kernel spend its time?  Builds a second copy of the
library with -DFSB_TIMELINE (thread 0 of every CTA stamps %globaltimer at: entry, ring primed,
first stage landed, main loop done, flush done, epilogue done), launches it ALONE on columns of
several sizes and prints, per size, the spread of every stamp relative to the first CTA's entry
plus the CUDA-event time of the launch.  Tool only; the product build has no probe code.

    python tools/timeline_probe.py [--build-only]        (JSON lines)
"""
import ctypes as C
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from libflagstats_b200 import build as B  # noqa: E402

SO = os.path.join(ROOT, "tools", "bin", "libflagstats_cuda_tl.so")


def build():
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    src = os.path.join(B.CSRC, "flagstat_capi.cu")
    deps = [os.path.join(B.CSRC, f) for f in os.listdir(B.CSRC)]
    if os.path.exists(SO) and all(os.path.getmtime(d) <= os.path.getmtime(SO) for d in deps):
        return
    subprocess.check_call([B.nvcc()] + B.NVCC_FLAGS + ["-DFSB_TIMELINE", "-DFSB_ALL_VARIANTS", "-o", SO, src])


def main():
    build()
    if "--build-only" in sys.argv:
        return
    import numpy as np
    import torch

    lib = C.CDLL(SO)
    lib.FLAGSTAT_cuda_device.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]
    lib.FLAGSTAT_cuda_synth_hiseqx.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, C.c_void_p]
    lib.FLAGSTAT_cuda_timeline_fetch.argtypes = [C.c_void_p, C.c_int]
    lib.FLAGSTAT_cuda_set_ctas_per_sm.argtypes = [C.c_int]
    big = 824_541_892
    data = torch.empty(big + 64, dtype=torch.int16, device="cuda")
    assert lib.FLAGSTAT_cuda_synth_hiseqx(data.data_ptr(), 0, big, 0, 0, None) == 0
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    out = torch.zeros(32, dtype=torch.int64, device="cuda")
    st = torch.cuda.Stream()
    names = ["entry", "ring_primed", "first_stage_landed", "loop_done", "flush_done", "epilogue_done"]

    def fetch():
        tl = np.zeros(8 * 2048, np.uint64)
        assert lib.FLAGSTAT_cuda_timeline_fetch(tl.ctypes.data, tl.size) == 0
        tl = tl.reshape(2048, 8)
        return tl[tl[:, 0] != 0].astype(np.int64)

    def summarize(t, rec):
        t0 = t[:, 0].min()
        for k, nm in enumerate(names):
            col = t[:, k][t[:, k] != 0] - t0
            if col.size:
                rec[nm] = {"min_us": round(col.min() / 1e3, 2), "median_us": round(float(np.median(col)) / 1e3, 2),
                           "max_us": round(col.max() / 1e3, 2)}
        # is the spread of loop_done systematic by SM?  mean loop_done per SM id, slowest / fastest five
        sm = t[:, 6] - 1
        done = (t[:, 3] - t0) / 1e3
        per = {}
        for a, b in zip(sm.tolist(), done.tolist()):
            per.setdefault(a, []).append(b)
        means = sorted((float(np.mean(v)), k) for k, v in per.items())
        rec["sms_used"] = len(per)
        rec["fastest_sms"] = [(k, round(m, 1)) for m, k in means[:5]]
        rec["slowest_sms"] = [(k, round(m, 1)) for m, k in means[-5:]]
        return rec

    lib.FLAGSTAT_cuda_set_dynamic.argtypes = [C.c_longlong, C.c_int]
    for per_sm, dyn in ((0, -1), (0, 1), (1, -1)):  # static split, dynamic claims (8 KiB chunks), static at 1 CTA / SM
        lib.FLAGSTAT_cuda_set_ctas_per_sm(per_sm)
        lib.FLAGSTAT_cuda_set_dynamic(dyn, 1)
        for n in (1 << 10, 1 << 20, 1 << 22, 1 << 24, 100_000_000, 1 << 28, big):
            if dyn == 1 and n < (1 << 22):
                continue
            # (a) one launch alone, column not in L2
            for rep in range(3):
                flush.fill_(rep)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                lib.FLAGSTAT_cuda_timeline_clear()
                e0.record(st)
                assert lib.FLAGSTAT_cuda_device(data.data_ptr(), n, out.data_ptr(), C.c_void_p(st.cuda_stream)) == 0
                e1.record(st)
                st.synchronize()
            rec = {"how": "alone", "split": "dynamic" if dyn == 1 else "static", "records": n, "ctas_per_sm": per_sm or "default",
                   "event_us": round(e0.elapsed_time(e1) * 1e3, 2), "ideal_us_at_7TBs": round(2 * n / 7.0e6, 2)}
            t = fetch()
            rec["ctas"] = int(t.shape[0])
            print(json.dumps(summarize(t, rec)), flush=True)
            # (b) the last of 8 back-to-back launches over rotating copies (steady state, from HBM)
            stride = (n + 15) // 8 * 8
            copies = max(1, min(8, (big - 8) // stride))
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            lib.FLAGSTAT_cuda_timeline_clear()
            e0.record(st)
            for i in range(8):
                assert lib.FLAGSTAT_cuda_device(data.data_ptr() + 2 * stride * (i % copies), n, out.data_ptr(),
                                                C.c_void_p(st.cuda_stream)) == 0
            e1.record(st)
            st.synchronize()
            rec = {"how": "last of 8 back-to-back", "split": "dynamic" if dyn == 1 else "static", "records": n, "ctas_per_sm": per_sm or "default",
                   "event_us_per_launch": round(e0.elapsed_time(e1) * 1e3 / 8, 2), "copies": copies,
                   "ideal_us_at_7TBs": round(2 * n / 7.0e6, 2)}
            t = fetch()
            rec["ctas"] = int(t.shape[0])
            print(json.dumps(summarize(t, rec)), flush=True)


if __name__ == "__main__":
    main()
