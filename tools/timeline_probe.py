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
    subprocess.check_call([B.nvcc()] + B.NVCC_FLAGS + ["-DFSB_TIMELINE", "-o", SO, src])


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
    for per_sm in (0, 1):
        lib.FLAGSTAT_cuda_set_ctas_per_sm(per_sm)
        for n in (1 << 10, 1 << 20, 1 << 22, 1 << 24, 100_000_000, 1 << 28, big):
            rows = []
            for rep in range(4):
                flush.fill_(rep)  # evict the column from L2
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                lib.FLAGSTAT_cuda_timeline_clear()
                with torch.cuda.stream(st):
                    e0.record(st)
                    assert lib.FLAGSTAT_cuda_device(data.data_ptr(), n, out.data_ptr(), C.c_void_p(st.cuda_stream)) == 0
                    e1.record(st)
                st.synchronize()
                tl = np.zeros(8 * 2048, np.uint64)
                assert lib.FLAGSTAT_cuda_timeline_fetch(tl.ctypes.data, tl.size) == 0
                tl = tl.reshape(2048, 8)
                live = tl[:, 0] != 0
                t = tl[live].astype(np.int64)
                t0 = t[:, 0].min()
                rec = {"records": n, "ctas_per_sm": per_sm or "default", "ctas": int(live.sum()), "rep": rep,
                       "event_us": round(e0.elapsed_time(e1) * 1e3, 2),
                       "ideal_us_at_7TBs": round(2 * n / 7.0e6, 2)}
                for k, nm in enumerate(names):
                    col = t[:, k][t[:, k] != 0] - t0
                    if col.size:
                        rec[nm] = {"min_us": round(col.min() / 1e3, 2), "median_us": round(float(np.median(col)) / 1e3, 2),
                                   "max_us": round(col.max() / 1e3, 2)}
                rows.append(rec)
            print(json.dumps(rows[-1]), flush=True)  # the last repetition (clocks settled)


if __name__ == "__main__":
    main()
