#!/usr/bin/env python3
"""Interleaved same-box A/B of kernel variants on the three workloads that matter:
HiSeqX-shaped (no QC-fail), uniform 12-bit (50 % QC-fail), HiSeqX with 1 % QC-fail.

    python tools/variant_ab.py 0 7 [more variants...] [s]

`s` = the samtools mode (exact n_pair_all, FLAGSTAT_cuda_samtools_device); its counters
are compared against the first variant's on every slot but 0 / 16.
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

PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])


def main():
    variants = [x if x == "s" else int(x) for x in sys.argv[1:]] or [0, 7]
    lib = fs.lib()
    N = synth.HISEQX_N
    data = {
        "hiseqx": synth.hiseqx_device(N),
        "uniform12": synth.uniform_device(N, 0, 0, 0x0FFF),
        "hiseqx_1pct_qcfail": synth.hiseqx_device(N, 0, 5, 10000),
    }
    out = torch.zeros(32, dtype=torch.int64, device="cuda")
    ms = C.c_float(0)
    ref = {}
    for rep in range(3):
        for name, t in data.items():
            for v in variants:
                mode = 2 if v == "s" else 0
                lib.FLAGSTAT_cuda_set_variant(0 if v == "s" else v)
                out.zero_()
                fs.check(lib.FLAGSTAT_cuda_time_device(t.data_ptr(), N, out.data_ptr(), 1, mode, C.byref(ms)), "t")
                got = out.cpu().tolist()
                if mode == 2:
                    got[0] = got[16] = 0
                ok = ref.setdefault(name, got) == got
                fs.check(lib.FLAGSTAT_cuda_time_device(t.data_ptr(), N, out.data_ptr(), 300, mode, C.byref(ms)), "t")
                gbs = 2 * N / (ms.value * 1e-3) / 1e9
                print(json.dumps({"workload": name, "variant": v, "rep": rep, "us": round(ms.value * 1e3, 2),
                                  "gbs": round(gbs, 1), "frac_of_measured_peak": round(gbs / PEAK, 3),
                                  "same_counters_as_first_variant": ok}), flush=True)
    lib.FLAGSTAT_cuda_set_variant(0)


if __name__ == "__main__":
    main()
