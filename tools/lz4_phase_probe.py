#!/usr/bin/env python3
"""Where does a warp of the LZ4 group decoder spend its cycles?  Builds a copy of the library with
-DFSB_LZ4_PROFILE (lane 0 of every warp sums clock64() deltas per phase of the group step), decodes
the bench containers (run-structured column, ratio ~5; i.i.d. column, ratio ~2.2) and prints the
share of each phase, cycles per group step and per sequence.  Tool only.

    python tools/lz4_phase_probe.py [n_blocks = 400]        (JSON lines)
"""
import ctypes as C
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "tools", "bin", "libflagstats_cuda_lz4prof.so")
os.environ["LIBFLAGSTATS_CUDA_SO"] = SO  # before the package is imported: _capi reads it at import time
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from libflagstats_b200 import build as B  # noqa: E402

PHASES = ["stage+sizes", "doubling+walk", "parse+scan", "literals", "parent pointers", "pointer jumping+root copy",
          "(sequences)", "(steps)", "slow-path sequences", "loop+flush"]


def build():
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    src = os.path.join(B.CSRC, "flagstat_capi.cu")
    deps = [os.path.join(B.CSRC, f) for f in os.listdir(B.CSRC)]
    if os.path.exists(SO) and all(os.path.getmtime(d) <= os.path.getmtime(SO) for d in deps):
        return
    subprocess.check_call([B.nvcc()] + B.NVCC_FLAGS + ["-DFSB_LZ4_PROFILE", "-o", SO, src])


def main():
    build()
    if "--build-only" in sys.argv:
        return
    os.environ["LIBFLAGSTATS_CUDA_SO"] = SO
    import numpy as np

    import containers
    import libflagstats_b200 as fs
    from libflagstats_b200 import blockfile

    n_blocks = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 400
    n = n_blocks * 512_000
    lib = fs.lib()
    lib.FLAGSTAT_cuda_lz4_profile_fetch.argtypes = [C.c_void_p, C.c_int]
    lib.FLAGSTAT_cuda_l4_profile_fetch.argtypes = [C.c_void_p, C.c_int]
    cta_phases = ["A1 successor tables + exit maps", "A2 chain", "A3 enumerate + parse + scan", "A4 descriptors + special",
                  "B0 tile index", "B1 sequences -> literals + parents (one thread per sequence)", "B1a staging",
                  "B3 pointer jumping", "B4 root -> byte", "B5 tile -> global"]
    for name, col in (("runs (ratio ~5)", containers.runs_column(n)), ("iid (ratio ~2.2)", containers.iid_column(n))):
        blob = containers.container(col, "lz4")
        # the CTA decoder (default): thread 0 of every CTA
        lib.FLAGSTAT_cuda_set_lz4_variant(2)
        blockfile.flagstat_container(blob, blockfile.LZ4)
        prof = np.zeros(16, np.uint64)
        lib.FLAGSTAT_cuda_l4_profile_fetch(prof.ctypes.data, 1)
        t0 = time.perf_counter()
        f, got = blockfile.flagstat_container(blob, blockfile.LZ4)
        dt = time.perf_counter() - t0
        lib.FLAGSTAT_cuda_l4_profile_fetch(prof.ctypes.data, 1)
        cyc = [int(x) for x in prof]
        tot = sum(cyc[:10]) + cyc[10] + cyc[14]
        print(json.dumps({"decoder": "cta", "b1_detail_cycles_per_tile": {"B1a staging": round(cyc[6] / max(cyc[13], 1)),
                          "B1c sequences (incl. waiting for the slowest thread)": round(cyc[5] / max(cyc[13], 1)),
                          "B1d long matches + loading the parents": round(cyc[14] / max(cyc[13], 1))}, "column": name, "blocks": n_blocks, "ratio": round(2 * n / len(blob), 2),
                          "container_call_s": round(dt, 4), "cycles_thread0_all_ctas": tot,
                          "share": {cta_phases[k]: round(cyc[k] / tot, 3) for k in range(10)},
                          "super_steps": cyc[12], "tiles": cyc[13], "jump_rounds": cyc[15],
                          "cycles_per_block": round(tot / n_blocks), "cycles_per_tile_phase_b": round((sum(cyc[5:10]) + cyc[10] + cyc[14]) / max(cyc[13], 1)),
                          "cycles_per_super_step_phase_a": round(sum(cyc[0:4]) / max(cyc[12], 1)),
                          "rounds_per_tile": round(cyc[15] / max(cyc[13], 1), 2)}), flush=True)
        lib.FLAGSTAT_cuda_set_lz4_variant(1)
        f, got = blockfile.flagstat_container(blob, blockfile.LZ4)  # warm (allocations)
        prof = np.zeros(16, np.uint64)
        lib.FLAGSTAT_cuda_lz4_profile_fetch(prof.ctypes.data, 1)
        t0 = time.perf_counter()
        f, got = blockfile.flagstat_container(blob, blockfile.LZ4)
        dt = time.perf_counter() - t0
        lib.FLAGSTAT_cuda_lz4_profile_fetch(prof.ctypes.data, 1)
        cyc = [int(x) for x in prof[:12]]
        tot = sum(cyc[k] for k in (0, 1, 2, 3, 4, 5, 8, 9))
        seqs = cyc[6]
        rec = {"decoder": "warp group", "column": name, "blocks": n_blocks, "ratio": round(2 * n / len(blob), 2), "records": got,
               "container_call_s": round(dt, 4), "sequences_in_group_steps": seqs, "group_steps": cyc[7],
               "slow_path_sequences": cyc[10], "sequences_per_step": round(seqs / max(cyc[7], 1), 1),
               "cycles_per_step": round(tot / max(cyc[7], 1), 0),
               "cycles_total_lane0_all_warps": tot,
               "share": {PHASES[k]: round(cyc[k] / tot, 3) for k in (0, 1, 2, 3, 4, 5, 8, 9)},
               "cycles_per_sequence": round(tot / max(seqs, 1), 1),
               "raw": cyc}
        print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
