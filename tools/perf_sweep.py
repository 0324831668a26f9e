#!/usr/bin/env python3
"""Device-resident kernel timing sweep (CUDA events inside the C ABI's
FLAGSTAT_cuda_time_device): workloads x lengths x kernel variants x grid sizes,
plus the 1,024,000-byte block streaming path.  Writes JSON lines to stdout.

    python tools/perf_sweep.py [--quick]
"""
import ctypes as C
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

PEAK = 6552.6
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def time_device(t, pospopcnt=False, iters=20):
    out = torch.zeros(32, dtype=torch.int64, device=t.device)
    ms = C.c_float(0)
    lib = fs.lib()
    fs.check(lib.FLAGSTAT_cuda_time_device(t.data_ptr(), t.numel(), out.data_ptr(), 3, int(pospopcnt),
                                           C.byref(ms)), "time")
    best = 1e30
    for _ in range(3):
        fs.check(lib.FLAGSTAT_cuda_time_device(t.data_ptr(), t.numel(), out.data_ptr(), iters,
                                               int(pospopcnt), C.byref(ms)), "time")
        best = min(best, ms.value)
    return best


def emit(**kw):
    print(json.dumps(kw), flush=True)


def main():
    quick = "--quick" in sys.argv
    lib = fs.lib()
    N = synth.HISEQX_N
    hiseqx = synth.hiseqx_device(N)
    uniform = synth.uniform_device(N, 0, 0, 0x0FFF)
    uniform16 = synth.uniform_device(N, 0, 1, 0xFFFF)
    fail1pct = synth.hiseqx_device(N, 0, 5, 10000)
    torch.cuda.synchronize()

    def row(name, t, **kw):
        ms = time_device(t, **kw)
        gbs = 2 * t.numel() / (ms * 1e-3) / 1e9
        emit(case=name, records=t.numel(), ms=ms, gbs=gbs, frac_of_measured_peak=gbs / PEAK,
             grec_s=t.numel() / (ms * 1e-3) / 1e9, **{k: v for k, v in kw.items() if k != "iters"})

    for variant in (0, 1, 2, 3, 5, 6):
        if lib.FLAGSTAT_cuda_set_variant(variant) < 0:
            continue  # A/B variant not compiled into this build (python -m libflagstats_b200.build --all-variants)
        row(f"hiseqx v{variant}", hiseqx)
        row(f"uniform12 (50% QC-fail) v{variant}", uniform)
        row(f"hiseqx 1% QC-fail v{variant}", fail1pct)
        row(f"pospopcnt uniform16 v{variant}", uniform16, pospopcnt=True)
        if variant in (0, 6):
            for per_sm in (1, 2, 3):
                lib.FLAGSTAT_cuda_set_ctas_per_sm(per_sm)
                row(f"hiseqx v{variant} ctas/sm={per_sm}", hiseqx)
            lib.FLAGSTAT_cuda_set_ctas_per_sm(0)
    lib.FLAGSTAT_cuda_set_variant(0)
    row("hiseqx base+1 record (unaligned)", hiseqx[1:])

    for per_sm in (1, 2, 3, 4):
        lib.FLAGSTAT_cuda_set_ctas_per_sm(per_sm)
        row(f"hiseqx ctas/sm={per_sm}", hiseqx)
        row(f"uniform12 ctas/sm={per_sm}", uniform)
    lib.FLAGSTAT_cuda_set_ctas_per_sm(0)

    # length sweep (BASELINE configs[2]); small sizes are launch-latency bound
    n = 1 << 10
    while n <= (1 << 29):
        for d in ((0,) if quick else (0, 1)):
            m = n + d * 127
            row(f"sweep hiseqx n={m}", hiseqx[:m], iters=50 if m < (1 << 24) else 20)
        n <<= 1 if quick else 1
        if quick:
            n <<= 1

    # 100 M uniform (BASELINE configs[0]); 200 MB is near L2 size -> rotate 4 distinct buffers
    bufs = [synth.uniform_device(100_000_000, i * 100_000_000, 0, 0x0FFF) for i in range(4)]
    out = torch.zeros(32, dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(2):
        for b in bufs:
            fs.flagstat_device(b, out=out)
    torch.cuda.synchronize()
    e0.record()
    reps = 5
    for _ in range(reps):
        for b in bufs:
            fs.flagstat_device(b, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / (reps * len(bufs))
    emit(case="inmemory 100M uniform12, 4 rotating buffers", records=100_000_000, ms=ms,
         gbs=0.2 / (ms * 1e-3), frac_of_measured_peak=0.2 / (ms * 1e-3) / PEAK)
    del bufs

    # streamed 1,024,000-byte blocks from pinned host (BASELINE configs[4])
    nblk = 400
    host = hiseqx[: nblk * fs.BLOCK_RECORDS].cpu().numpy().view(np.uint16)
    for slots in (2, 3, 4, 8):
        with fs.BlockStream(0, fs.BLOCK_RECORDS, slots) as bs:
            for rep in range(2):
                t0 = time.perf_counter()
                for i in range(nblk):
                    slot = bs.acquire()
                    slot[:] = host[i * fs.BLOCK_RECORDS:(i + 1) * fs.BLOCK_RECORDS]
                    bs.submit(fs.BLOCK_RECORDS)
                f = bs.finish()
                t1 = time.perf_counter()
            with_fill = t1 - t0
            # producer already has the data in the pinned slots: submit only
            t0 = time.perf_counter()
            for i in range(nblk):
                bs.acquire()
                bs.submit(fs.BLOCK_RECORDS)
            bs.finish()
            t1 = time.perf_counter()
            only = t1 - t0
        emit(case=f"stream 1,024,000-B blocks x{nblk}, slots={slots}",
             gbs_with_host_memcpy=nblk * 1.024e-3 / with_fill, gbs_submit_only=nblk * 1.024e-3 / only,
             grec_s_submit_only=nblk * fs.BLOCK_RECORDS / only / 1e9, total=int(f[9] + f[25]))


if __name__ == "__main__":
    main()
