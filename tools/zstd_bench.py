#!/usr/bin/env python3
"""Zstd containers through the GPU, both decoders (two-stage default; one thread per frame, FLAGSTAT_CUDA_ZSTD_VARIANT=0):
end-to-end time of FLAGSTAT_cuda_container_u64 for the two bench columns (run-structured, i.i.d.) at 400 and 1600
frames, libzstd level 1, next to the reference's block loop on this host's cores.  JSON lines.
    python tools/zstd_bench.py [--quick]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import containers  # noqa: E402
import libflagstats_b200 as fs  # noqa: E402
from libflagstats_b200 import blockfile  # noqa: E402


def main():
    quick = "--quick" in sys.argv
    only_default = "--only-default" in sys.argv  # the default decoder only, no CPU reference (parameter sweeps)
    threads = len(os.sched_getaffinity(0))
    for n_blocks in ((400,) if quick else (400, 1600)):
        n = n_blocks * 512_000 + 12_345
        for name, col in (("runs", containers.runs_column(n)), ("iid", containers.iid_column(n))):
            blob = containers.container(col, "zstd", 1)
            want = None
            for variant in (("1",) if only_default else ("1", "0")):
                if variant == "0" and n_blocks > 400:
                    continue
                os.environ["FLAGSTAT_CUDA_ZSTD_VARIANT"] = variant
                best = 1e30
                for _ in range(4 if variant == "1" else 2):
                    t0 = time.perf_counter()
                    f, got = blockfile.flagstat_container(blob, blockfile.ZSTD)
                    best = min(best, time.perf_counter() - t0)
                want = want if want is not None else f.tolist()
                print(json.dumps({"column": name, "frames": n_blocks + 1, "ratio": round(2 * n / len(blob), 2),
                                  "variant": "two-stage" if variant == "1" else "one thread per frame",
                                  "container_call_ms": round(best * 1e3, 3), "gbs_records": round(2 * n / best / 1e9, 2),
                                  "same_counters": f.tolist() == want and got == n}), flush=True)
            os.environ["FLAGSTAT_CUDA_ZSTD_VARIANT"] = "1"
            if only_default:
                continue
            try:
                from oracle import oracle as O  # the CPU side only
                if O.ref_container_available("zstd"):
                    fa, na, sa, _ = O.ref_container_mt(blob, "zstd", threads)
                    sa = min([sa] + [O.ref_container_mt(blob, "zstd", threads)[2] for _ in range(2)])
                    core19 = [i for i in fs.CORE20 if i != 9]
                    print(json.dumps({"column": name, "frames": n_blocks + 1, "variant": f"reference loop, {threads} host threads",
                                      "container_call_ms": round(sa * 1e3, 3), "gbs_records": round(2 * na / sa / 1e9, 2),
                                      "same_counters": [int(fa[i]) for i in core19] == [want[i] for i in core19]}), flush=True)
            except Exception as exc:
                print(json.dumps({"cpu_reference_error": repr(exc)}), flush=True)


if __name__ == "__main__":
    main()
