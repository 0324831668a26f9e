#!/usr/bin/env python3
"""Length sweep of the shipping kernel (BASELINE.json configs[2] and configs[0]).

    python tools/length_sweep.py [--quick] > profiles/<tag>_length_sweep.jsonl

n = 2^10 ... 2^32 records (and n-1, n+1, n+127, n+511), bases misaligned by 0/1/3/7 records,
flagstat on HiSeqX-shaped and U(0,4095) columns, raw pospopcnt on U(0,65535).  Every point is
timed with CUDA events inside the C ABI (FLAGSTAT_cuda_time_device_rot): back-to-back launches
that ROTATE over distinct copies of the column covering >= 1 GiB where the column itself is
smaller than that, so that mid-size columns are read from HBM and not from the 126 MB L2
(`hot` rows repeat the small sizes over ONE copy: the L2-resident figure).  The 100 M-record
U(0,4095) rows are BASELINE configs[0] on the GPU.  JSON lines on stdout.
"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import libflagstats_b200 as fs  # noqa: E402
from libflagstats_b200 import synth  # noqa: E402

try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    PEAK = 6650.0

MODES = {"flagstat": 0, "pospopcnt": 1, "samtools": 2}


def emit(**kw):
    print(json.dumps(kw), flush=True)


def timed(region, n, base, mode, rotate=True, target_ms=60.0, overlapped=False):
    """ms per launch over `region` (a device tensor of 16-bit words)."""
    lib = fs.lib()
    out = torch.zeros(32, dtype=torch.int64, device=region.device)
    stride = (n + base + 7 + 8) // 8 * 8  # copies 16-byte aligned relative to each other
    cap = (region.numel() - base) // stride if stride else 1
    if rotate:
        want = max(1, -(-(1 << 29) // max(stride, 1)))  # >= 1 GiB of distinct bytes
        n_rot = max(1, min(cap, want, 4096))
    else:
        n_rot = 1
    ms = C.c_float(0)
    mode = mode | (4 if overlapped else 0)
    ptr = region.data_ptr() + 2 * base
    fs.check(lib.FLAGSTAT_cuda_time_device_rot(ptr, n, stride, n_rot, out.data_ptr(), max(3, min(n_rot, 50)),
                                               mode, C.byref(ms)), "time")
    est = max(ms.value, 1e-3)
    iters = int(max(10, min(3000, target_ms / est)))
    best = 1e30
    for _ in range(3):
        fs.check(lib.FLAGSTAT_cuda_time_device_rot(ptr, n, stride, n_rot, out.data_ptr(), iters, mode,
                                                   C.byref(ms)), "time")
        best = min(best, ms.value)
    return best, n_rot, iters


def row(case, region, n, base=0, mode="flagstat", rotate=True, overlapped=False):
    ms, n_rot, iters = timed(region, n, base, MODES[mode], rotate, overlapped=overlapped)
    gbs = 2 * n / (ms * 1e-3) / 1e9
    emit(case=case, mode=mode, records=n, base_offset_records=base, ms=round(ms, 6), us=round(ms * 1e3, 3),
         gbs=round(gbs, 1), frac_of_measured_peak=round(gbs / PEAK, 4), grec_s=round(n / (ms * 1e-3) / 1e9, 3),
         copies_rotated=n_rot, distinct_bytes=2 * n * n_rot, from_hbm=bool(2 * n * n_rot > (256 << 20)),
         launches="overlapped (programmatic dependent launch)" if overlapped else "serialised in one stream",
         iters=iters)


def main():
    quick = "--quick" in sys.argv
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    big = (1 << 32) + 1024  # records: the 2^32 point needs the u64-length entry (legacy max is 2^32 - 1)
    hiseqx = synth.hiseqx_device(big, 0, device=dev)
    small = (1 << 30) + 1024
    uniform = synth.uniform_device(small, 0, 0, 0x0FFF, device=dev)
    uniform16 = synth.uniform_device(small, 0, 1, 0xFFFF, device=dev)
    torch.cuda.synchronize()
    emit(kernel=fs.lib().FLAGSTAT_cuda_kernel_name(0).decode(), peak_gbs=PEAK, gpu=torch.cuda.get_device_name(0))

    exps = range(10, 33, 2 if quick else 1)
    for e in exps:
        n = 1 << e
        row("sweep hiseqx", hiseqx, n)
        row("sweep hiseqx, overlapped launches", hiseqx, n, overlapped=True)
        if n <= (1 << 25):
            row("sweep hiseqx hot (one copy, L2-resident)", hiseqx, n, rotate=False)
        if not quick or e % 4 == 0:
            for d in (-1, 1, 127, 511):
                if n + d <= hiseqx.numel():
                    row(f"sweep hiseqx n{d:+d}", hiseqx, n + d)
            for base in (1, 3, 7):
                if n + base <= hiseqx.numel() - 16:
                    row(f"sweep hiseqx base+{base}", hiseqx, n, base=base)
        if n <= (1 << 30):
            row("sweep uniform12", uniform, n)
            row("sweep pospopcnt uniform16", uniform16, n, mode="pospopcnt")
            if not quick and e % 4 == 0:
                row("sweep samtools uniform12", uniform, n, mode="samtools")
    # BASELINE configs[0]: 100 M records U(0,4095); and the same length HiSeqX-shaped
    for n in (100_000_000, 16_777_216, 50_000_000):
        for ov in (False, True):
            tag = ", overlapped launches" if ov else ""
            row(f"inmemory {n} uniform12 (configs[0] shape){tag}", uniform, n, overlapped=ov)
            row(f"inmemory {n} hiseqx{tag}", hiseqx, n, overlapped=ov)
            row(f"inmemory {n} pospopcnt uniform16{tag}", uniform16, n, mode="pospopcnt", overlapped=ov)
    # static split vs dynamic claims (8 / 16 KiB chunks), and where the default threshold should sit
    for label, minc, cg in (("static", -1, 1), ("dynamic 8 KiB", 1, 1), ("dynamic 16 KiB", 1, 2)):
        if fs.lib().FLAGSTAT_cuda_set_dynamic(minc, cg) != 0:
            continue  # the dynamic kernel is in -DFSB_ALL_VARIANTS builds only (LIBFLAGSTATS_CUDA_SO=tools/bin/...)
        for n in (1 << 22, 1 << 23, 1 << 24, 1 << 25, 50_000_000, 100_000_000, 1 << 28, 824_541_892, 1 << 31):
            row(f"split={label} hiseqx", hiseqx, n)
            if n in (100_000_000, 824_541_892):
                row(f"split={label} uniform12", uniform, n)
                row(f"split={label} hiseqx, overlapped launches", hiseqx, n, overlapped=True)
    fs.lib().FLAGSTAT_cuda_set_dynamic(-1, 0)
    # grid-size sensitivity in the mid-size regime
    for per_sm in (1, 2):
        fs.lib().FLAGSTAT_cuda_set_ctas_per_sm(per_sm)
        for n in (1 << 22, 1 << 24, 1 << 25, 50_000_000, 1 << 26, 100_000_000, 1 << 27, 200_000_000, 1 << 28,
                  400_000_000, 1 << 29, 824_541_892):
            row(f"ctas_per_sm={per_sm} hiseqx", hiseqx, n)
            if n >= 50_000_000 and n <= (1 << 30):
                row(f"ctas_per_sm={per_sm} uniform12", uniform, n)
                row(f"ctas_per_sm={per_sm} hiseqx, overlapped launches", hiseqx, n, overlapped=True)
    fs.lib().FLAGSTAT_cuda_set_ctas_per_sm(0)


if __name__ == "__main__":
    main()
