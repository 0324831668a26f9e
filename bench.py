#!/usr/bin/env python3
"""bench.py -- the flagstat hot path on N B200s (contract: see DESIGN.md section 6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one pass of the hot path over one batch of synthetic input: every
rank runs the sm_100a kernel over ITS shard of the FLAG column (already
resident in HBM); the same kernel launch exchanges the 32 counters with the
other ranks through peer-mapped memory (--exchange nccl: a separate NCCL
all-reduce instead).  Workload at
any N: BASELINE.json configs[1] per GPU -- 824,541,892 HiSeqX-shaped records
(1.65 GB, > L2) per rank, rank r holding global records [r*n, (r+1)*n) of the
periodic generator, so the exact global answer is N x KAT-E (weak scaling).

Prints ONE JSON line on rank 0.  `value` is device-resident whole-job
records/s; `e2e` is the same metric through the public host-pointer API
(FLAGSTAT_cuda_u64 on pinned host memory: H2D of the whole shard + D2H of the
counters inside the timed region); `roofline` is the kernel's algorithmic
bytes / CUDA-event time against the measured HBM peak; `cpu_baseline` is the
unmodified reference timed on this box's host cores on a bounded sample.

--impl reference times the reference's own CPU implementation (oracle/_ref,
FLAGSTAT_avx512 or whatever FLAGSTATS_get_function picks here) with all host
threads on the same workload definition; rank 0 only.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

HISEQX_N = 824_541_892
METRIC = "flag_records_per_s"
UNIT = "records/s"


def workload_config(n_gpus: int, per_gpu: int) -> dict:
    return {
        "workload": f"hiseqx_shaped_{per_gpu}_records_per_gpu (BASELINE configs[1])",
        "records_per_gpu": per_gpu,
        "bytes_per_gpu": 2 * per_gpu,
        "global_records": per_gpu * n_gpus,
        "cache": "input (1.65 GB/GPU) larger than L2 (126 MB); no flush needed",
        "parallelism": f"range-shard x{n_gpus} + sum of 32 x u64 counters across ranks",
    }


def load_peaks() -> tuple:
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json, copy read+write)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def load_traffic():
    """Per-launch DRAM bytes of the flagstat kernel from the committed ncu
    capture (profiles/ncu_traffic.json), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as fh:
            return json.load(fh)
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-i", str(self.index), "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def wait_first_sample(self, timeout_s: float) -> None:
        t0 = time.perf_counter()
        while self.proc is not None and not self.rows and time.perf_counter() - t0 < timeout_s:
            time.sleep(0.01)

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def stop(self, t0: float, t1: float) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        rows = [r for t, r in self.rows if t0 <= t <= t1 + 0.06] or [r for _, r in self.rows]
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                pw.append(float(r[3]))
                for nme, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except (ValueError, IndexError):
                continue
        return {
            "sm_mhz": statistics.median(sm) if sm else None,
            "sm_max_mhz": max(mx) if mx else None,
            "power_w_max": max(pw) if pw else None,
            "samples": len(sm),
            "reasons": sorted(reasons),
        }


# ---------------------------------------------------------------------------
# reference arm
# ---------------------------------------------------------------------------
def cpu_reference_run(sample_records: int, steps: int, warmup: int, threads: int):
    """Time the unmodified reference (or, if its prebuilt shim is missing, the
    oracle port) on `sample_records` HiSeqX-shaped records."""
    from oracle import oracle as O

    a = O.synth_hiseqx(0, sample_records)
    if O.reference() is not None:
        kind, kernel = "reference", O.best_reference_kernel()

        def run(nt):
            f, sec = O.ref_flagstat_mt(kernel, a, nt)
            return f, sec
    else:
        kind, kernel = "port", "oracle_flagstat_simd_u64"

        def run(nt):
            t = time.perf_counter()
            f = O.flagstat_simd(a)
            return f, time.perf_counter() - t
        threads = 1
    for _ in range(warmup):
        run(threads)
    times = []
    f = None
    for _ in range(steps):
        f, sec = run(threads)
        times.append(sec)
    one = min(run(1)[1] for _ in range(3))
    want = O.numpy_flagstat(a[: 1 << 22])
    got = O.ref_flagstat_mt(kernel, a[: 1 << 22], threads)[0] if kind == "reference" else want
    ok = all(int(got[i]) == int(want[i]) for i in O.CORE20)
    return {
        "kind": kind, "kernel": kernel, "threads": threads, "times": times,
        "sample_records": sample_records, "one_thread_s": one, "verified": ok,
    }


def cpu_baseline_set(threads: int, n: int = 100_000_000) -> dict:
    """BASELINE.md section 4 / benchmark/inmemory.cpp:59-102,108-116: the reference's kernels on
    n records U(0,4095), each at 1 thread (what the reference ships) and on all host cores
    (our pthread range wrapper around the unmodified kernels).  Best of a few runs, steady_clock
    inside the shim (the reference's own timer wraps at 65 ms, inmemory.cpp:134)."""
    from oracle import oracle as O

    if O.reference() is None:
        return {"unavailable": "oracle/_ref was not prebuilt"}
    a = O.synth_uniform(0, n, 0, 0x0FFF)
    want = None
    out = {"records": n, "input": "U(0,4095) index hash (benchmark/generate.cpp:11 distribution)", "threads_all": threads,
           "kernels": {}}
    runnable = set(O.ref_kernels())
    for k in ("scalar", "avx2", "avx512", "avx512_improved3"):
        if k not in runnable:
            out["kernels"][k] = None
            continue
        m1 = n // 10 if k == "scalar" else n  # the branchy scalar loop takes ~25 ns/record on uniform flags
        one = min(O.ref_flagstat_mt(k, a[:m1], 1)[1] for _ in range(2 if k == "scalar" else 3))
        f, _ = O.ref_flagstat_mt(k, a, threads)
        alls = min(O.ref_flagstat_mt(k, a, threads)[1] for _ in range(4))
        if want is None:
            want = [int(f[i]) for i in O.CORE20 if i != 9]
        out["kernels"][k] = {
            "one_thread_grec_s": m1 / one / 1e9, "one_thread_records": m1,
            "all_cores_grec_s": n / alls / 1e9, "all_cores_gbs": 2 * n / alls / 1e9,
            "agrees_with_scalar_on_core19": [int(f[i]) for i in O.CORE20 if i != 9] == want,
        }
    if hasattr(O.reference(), "ref_pospopcnt_mt"):
        one = min(O.ref_pospopcnt_mt(a, 1)[1] for _ in range(3))
        alls = min(O.ref_pospopcnt_mt(a, threads)[1] for _ in range(4))
        out["kernels"]["STORM_pospopcnt_u16"] = {"one_thread_grec_s": n / one / 1e9, "one_thread_records": n,
                                                 "all_cores_grec_s": n / alls / 1e9, "all_cores_gbs": 2 * n / alls / 1e9}
    out["dispatched_by_FLAGSTATS_get_function"] = O.best_reference_kernel()
    return out


def cpu_file_reference(col, blobs: dict, threads: int) -> dict:
    """The reference's readers of its own FLAG files on this host (benchmark/flagstats.cpp:288-358 LZ4,
    :636-676 Zstd, :415-468 raw): the system codec per block, then the kernel
    FLAGSTATS_get_function(N) returns -- as shipped (1 thread, timed on the first blocks only) and with
    the blocks dealt to all cores.  Returns per file {counters, grec_s_1thread, grec_s_all_cores}."""
    from oracle import oracle as O

    res = {}
    if O.reference() is None or not hasattr(O.reference(), "ref_container_mt"):
        return res
    kernel = O.best_reference_kernel()
    sample = col[: min(col.size, 100 * 512_000)]
    one = min(O.ref_flagstat_mt(kernel, sample, 1)[1] for _ in range(2))
    f, _ = O.ref_flagstat_mt(kernel, col, threads)
    alls = min(O.ref_flagstat_mt(kernel, col, threads)[1] for _ in range(3))
    res["raw"] = {"counters": [int(x) for x in f], "grec_s_1thread": sample.size / one / 1e9,
                  "grec_s_all_cores": col.size / alls / 1e9, "kernel": kernel,
                  "note": "kernel over the column in memory (the file read of flagstats.cpp:415-468 not included)"}
    for name, (codec, blob) in blobs.items():
        if not O.ref_container_available(codec):
            continue
        head = bytes(blob[: _container_prefix(blob, 100)])
        f1, n1, s1, d1 = O.ref_container_mt(head, codec, 1)
        fa, na, sa, da = O.ref_container_mt(blob, codec, threads)
        sa = min([sa] + [O.ref_container_mt(blob, codec, threads)[2] for _ in range(2)])
        res[name] = {"counters": [int(x) for x in fa], "grec_s_1thread": n1 / s1 / 1e9,
                     "decode_share_1thread": d1 / s1, "grec_s_all_cores": na / sa / 1e9, "kernel": kernel}
    return res


def _container_prefix(blob, n_blocks: int) -> int:
    """Byte length of the first n_blocks [int32 raw][int32 comp][payload] records."""
    import struct
    pos = 0
    for _ in range(n_blocks):
        if pos + 8 > len(blob):
            break
        _raw, comp = struct.unpack_from("<ii", blob, pos)
        pos += 8 + comp
    return min(pos, len(blob))


def cpu_dropin_loop() -> dict:
    """oracle/_ref/dropin_check --time: the reference's block loop (benchmark/flagstats.cpp:304,328-329)
    through its own dispatcher compiled with integration/libflagstats_h_cuda.patch -- pageable
    512,000-record blocks into FLAGSTAT_cuda vs the reference's CPU kernel at 1 thread, and the
    per-call crossover length."""
    exe = os.path.join(ROOT, "oracle", "_ref", "dropin_check")
    if not os.path.exists(exe):
        return {"unavailable": "oracle/_ref/dropin_check was not prebuilt"}
    try:
        p = subprocess.run([exe, "--time"], capture_output=True, text=True, timeout=240)
    except Exception as exc:
        return {"error": repr(exc)}
    rows = []
    for line in p.stdout.splitlines():
        try:
            rows.append(json.loads(line))
        except ValueError:
            pass
    out = {"rc": p.returncode, "min_len_default": next((r["cuda_min_len_default"] for r in rows if "cuda_min_len_default" in r), None)}
    for r in rows:
        if "loop" in r:
            out[r["loop"]] = {k: v for k, v in r.items() if k != "loop"}
        if "crossover_records" in r:
            out["crossover_records_pageable"] = r["crossover_records"]
    out["per_call_us_pageable"] = {str(r["n"]): [r["cuda_us"], r["cpu_us"]] for r in rows if r.get("sweep") == "pageable"}
    return out


def cpu_model() -> str:
    try:
        with open("/proc/cpuinfo") as fh:
            for line in fh:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def reference_arm(args) -> int:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    sample = min(HISEQX_N, args.cpu_sample)
    r = cpu_reference_run(sample, args.steps, args.warmup, threads)
    sec = min(r["times"])  # best step: the generous reading for the baseline on a noisy host
    value = sample / sec
    line = {
        "impl": "reference",
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u16", "data": "synthetic",
        "config": workload_config(args.gpus, HISEQX_N),
        "cpu_baseline": {
            "value": value, "unit": UNIT, "cores": r["threads"], "kind": r["kind"],
            "sample": f"first {sample} records of the workload per step; kernel {r['kernel']} "
                      f"(FLAGSTATS_get_function's choice on this CPU), range-sharded over "
                      f"{r['threads']} pthreads; best of {args.steps} steps (median "
                      f"{statistics.median(r['times']) * 1e3:.2f} ms)",
            "one_thread_value": sample / r["one_thread_s"], "cpu": cpu_model(),
            "verified": r["verified"],
        },
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------
STRONG_N = 1 << 34


def strong_scaling_leg(args, fs, sharded, synth, xchg, overlap, dev, stream, world, rank, fence, dist, torch, np) -> dict:
    """BASELINE.json configs[3]: a FIXED 2^34-record column (the reference's uint32_t len cannot
    even express it, libflagstats.h:170) split into `world` contiguous ranges."""
    lo, hi = sharded.shard_range(STRONG_N, world, rank)
    shard = synth.hiseqx_device(hi - lo, start=lo, device=dev)
    out = torch.zeros(32, dtype=torch.int64, device=dev)
    torch.cuda.synchronize(dev)
    deferred = xchg is not None and overlap and args.deferred

    def run(k, d):
        for _ in range(k):
            if xchg is not None:
                xchg.flagstat(shard, out=out, accumulate=False, stream=stream, deferred=d)
            else:
                out.zero_()
                fs.flagstat_device(shard, out=out, stream=stream)
                sharded.allreduce_counters(out)
        if xchg is not None and d:
            xchg.collect(stream=stream)

    def timed(k, d):
        run(3, d)
        fence()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        run(k, d)
        b.record(stream)
        fence()
        t = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / k, out.cpu().numpy().view(np.uint64).tolist()

    k = max(3, min(args.steps, 20))
    ms, got = timed(k, deferred)
    ms_serial, got_serial = None, None
    if xchg is not None and overlap:
        xchg.set_overlap(False)
        ms_serial, got_serial = timed(k, False)
        xchg.set_overlap(True)
    with open(os.path.join(ROOT, "tests", "golden", "flagstat_golden.json")) as fh:
        kat = json.load(fh)["kat_16g"]
    assert kat["spec"]["n"] == STRONG_N
    verified = got == kat["cuda_expected"] and (got_serial is None or got_serial == kat["cuda_expected"])
    hint = None
    try:  # strong-scaling efficiency against the committed 1-GPU measurement of the same leg (a hint:
        # the driver computes its own from the per-N lines)
        with open(os.path.join(ROOT, "profiles", "strong_2p34_n1.json")) as fh:
            n1 = json.load(fh)
        hint = (STRONG_N / (ms * 1e-3)) / (world * n1["value"])
    except Exception:
        pass
    del shard
    torch.cuda.empty_cache()
    return {
        "workload": "BASELINE configs[3]: 2^34 HiSeqX-shaped records range-sharded over the ranks (strong scaling)",
        "records": STRONG_N, "records_this_rank": hi - lo, "bytes_per_gpu": 2 * (hi - lo), "steps": k,
        "ms_per_step": ms, "value": STRONG_N / (ms * 1e-3), "unit": UNIT,
        "ms_per_step_serialised_launches": ms_serial,
        "value_serialised": (STRONG_N / (ms_serial * 1e-3)) if ms_serial else None,
        "gbs_per_gpu": 2 * (hi - lo) / (ms * 1e-3) / 1e9,
        "verified": bool(verified), "efficiency_vs_n1_hint": hint,
        "steps_overlapped": overlap, "deferred_collection": deferred,
    }


def inmemory_leg(fs, synth, torch, dev, peak) -> dict:
    """BASELINE configs[0] on the GPU: 100 M records U(0,4095) (benchmark/inmemory.cpp:108-116 with
    size 100 M), device-resident.  200 MB is close to the 126 MB L2, so launches rotate over 5
    distinct copies (1 GB); CUDA events inside the C ABI (FLAGSTAT_cuda_time_device_rot)."""
    import ctypes as C

    n, copies = 100_000_000, 5
    stride = (n + 8 + 7) // 8 * 8
    out = torch.zeros(32, dtype=torch.int64, device=dev)
    res = {"records": n, "copies_rotated": copies, "timing": "CUDA events around back-to-back launches, mean"}
    for name, gen, mode in (("flagstat_uniform12", "u12", 0), ("flagstat_hiseqx_shaped", "hx", 0),
                            ("pospopcnt_uniform16", "u16", 1), ("samtools_uniform12", "u12", 2)):
        buf = torch.empty(stride * copies, dtype=torch.int16, device=dev)
        for c in range(copies):
            view = buf[c * stride: c * stride + n]
            if gen == "hx":
                synth.hiseqx_device(n, start=c * n, device=dev, out=view)
            else:
                synth.uniform_device(n, c * n, 0 if gen == "u12" else 1, 0x0FFF if gen == "u12" else 0xFFFF,
                                     device=dev, out=view)
        torch.cuda.synchronize(dev)
        ms = C.c_float(0)
        res[name] = {}
        # one call at a time (every launch waits for the previous one), and back-to-back calls that
        # overlap head to tail (FLAGSTAT_cuda_device_overlapped: hides the ~10 us of launch gap, ramp
        # and tail a 40 us launch carries)
        for key, m in (("serialised_launches", mode), ("overlapped_launches", mode | 4)):
            fs.check(fs.lib().FLAGSTAT_cuda_time_device_rot(buf.data_ptr(), n, stride, copies, out.data_ptr(), 50, m,
                                                            C.byref(ms)), "time_device_rot")
            best = 1e30
            for _ in range(3):
                fs.check(fs.lib().FLAGSTAT_cuda_time_device_rot(buf.data_ptr(), n, stride, copies, out.data_ptr(), 500, m,
                                                                C.byref(ms)), "time_device_rot")
                best = min(best, ms.value)
            gbs = 2 * n / (best * 1e-3) / 1e9
            res[name][key] = {"us_per_launch": best * 1e3, "grec_s": n / (best * 1e-3) / 1e9, "gbs": gbs,
                              "frac_of_measured_peak": gbs / peak}
        del buf
    torch.cuda.empty_cache()
    return res


def file_leg(fs, torch, dev, threads: int, pcie_gbs: float) -> dict:
    """the reference's FLAG files end to end on the GPU (file in the page cache ->
    counters) next to the reference's own readers of the same files on this host's cores
    (cpu_file_reference).  400 blocks of 1,024,000 bytes (204.8 M records): raw .bin, LZ4 containers of
    a run-structured column (ratio ~5) and of an i.i.d. one (ratio ~2.2), a Zstd level-1 container."""
    import tempfile

    import numpy as np

    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import containers
    from libflagstats_b200 import blockfile

    n = 400 * 512_000 + 12_345
    cols = {"runs": containers.runs_column(n), "iid": containers.iid_column(n)}
    blobs = {"lz4_runs": ("lz4", containers.container(cols["runs"], "lz4")),
             "lz4_iid": ("lz4", containers.container(cols["iid"], "lz4"))}
    if containers.libzstd() is not None:
        blobs["zstd1_runs"] = ("zstd", containers.container(cols["runs"], "zstd", 1))
    tmp = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    res = {"records": n, "blocks": 401, "pcie_h2d_probe_gbs": pcie_gbs, "files": {}}
    try:
        cpu = {"runs": cpu_file_reference(cols["runs"], {k: v for k, v in blobs.items() if k.endswith("runs")}, threads),
               "iid": cpu_file_reference(cols["iid"], {k: v for k, v in blobs.items() if k.endswith("iid")}, threads)}
        jobs = [("raw_bin", "runs", None, ".bin")] + [(k, k.split("_")[-1], v[1], ".lz4" if v[0] == "lz4" else ".zst")
                                                      for k, v in blobs.items()]
        core19 = [i for i in fs.CORE20 if i != 9]
        for name, colname, blob, ext in jobs:
            path = os.path.join(tmp, name + ext)
            if blob is None:
                cols[colname].tofile(path)
            else:
                with open(path, "wb") as fh:
                    fh.write(blob)
            size = os.path.getsize(path)
            best = 1e30
            for _ in range(4):
                t0 = time.perf_counter()
                f, got_n = blockfile.flagstat_file(path)
                best = min(best, time.perf_counter() - t0)
            ref = cpu[colname].get("raw" if blob is None else name)
            ok = got_n == n and int(f[9]) + int(f[25]) == n
            if ref is not None:  # slot 9 is left out: the reference's scalar tail kernel never writes it (libflagstats.h:127)
                ok = ok and [int(f[i]) for i in core19] == [ref["counters"][i] for i in core19]
            res["files"][name] = {
                "file_bytes": size, "ratio": 2 * n / size, "seconds": best, "grec_s": n / best / 1e9,
                "gbs_records": 2 * n / best / 1e9, "gbs_file": size / best / 1e9, "verified": bool(ok),
                "cpu_reference": ({k: v for k, v in ref.items() if k != "counters"} if ref else None),
                "speedup_vs_reference_1thread": (n / best / 1e9 / ref["grec_s_1thread"]) if ref else None,
                "speedup_vs_reference_all_cores": (n / best / 1e9 / ref["grec_s_all_cores"]) if ref else None,
            }
            os.remove(path)
    finally:
        try:
            os.rmdir(tmp)
        except OSError:
            pass
    return res


def ours(args) -> int:
    import numpy as np
    import torch
    import torch.distributed as dist

    import libflagstats_b200 as fs
    from libflagstats_b200 import sharded, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback; use --impl reference)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert fs.available() > 0

    n = args.records
    start = rank * n
    data = synth.hiseqx_device(n, start=start, device=dev)
    counters = torch.zeros(32, dtype=torch.int64, device=dev)
    # all steps run on one dedicated (non-default) stream: overlapped launches need a stream
    # that takes the programmatic-serialization attribute, and the input is complete and
    # visible before the first step (synchronised here)
    torch.cuda.synchronize(dev)
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)

    overlap = args.exchange == "fused" and not args.no_overlap
    xchg = sharded.FusedExchange(device=dev, overlap=overlap) if args.exchange == "fused" else None

    def step_nccl():
        counters.zero_()
        fs.flagstat_device(data, out=counters, stream=stream)
        sharded.allreduce_counters(counters)

    deferred = xchg is not None and overlap and args.deferred

    def step_fused():
        # ONE kernel launch per rank: count the shard, push the 32 totals into every
        # peer's exchange buffer over NVLink, write the global counters (overwrite mode: no
        # memset either).  By default the launch's last CTA also waits for the peers' totals of
        # this step (they arrive while the NEXT step's CTAs are already streaming: overlapped
        # launches).  --deferred: the launch only pushes, the next step's launch collects (and the
        # last step's are collected by xchg.collect() inside the timed region).  Both orders are
        # measured in every multi-GPU run; they tie at N = 2, 4 and 8 (the headline leg, measured
        # right after the settle loop, is the ~1 % slower one at N = 8 whichever order it uses:
        # profiles/r7b_*, r9e_*, r9f_*), so the simpler one leads.
        xchg.flagstat(data, out=counters, accumulate=False, stream=stream, deferred=deferred)

    def finish_steps():
        if deferred:
            xchg.collect(stream=stream)

    step = step_fused if xchg is not None else step_nccl

    def fence():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    # nvidia-smi starts polling NOW, not right in front of the timed region: while it initialises
    # (it attaches to every GPU of the box: longer the more GPUs there are) kernel launches on all
    # GPUs are disturbed -- the first timed leg of a multi-GPU run used to come out 3 - 7 us per step
    # slower than the same steps measured a little later (profiles/r4z_*)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()

    # warm-up: W steps, then keep stepping until ~0.3 s of load has passed so that the
    # timed region sees the clocks the GPU sustains under this kernel (on this pool the
    # first ~100 ms after idle run ~5 % faster than the power-capped steady state)
    for _ in range(max(args.warmup, 3)):
        step()
    fence()
    # nvidia-smi must have reached its steady polling state BEFORE the settle loop, not between it and
    # the timed region (on an 8-GPU box its start-up takes seconds, during which the GPUs would sit
    # idle): the K timed steps follow the settle loop directly and see the sustained, power-capped
    # clocks.  (The comparison legs further down run 20 steps each after short host-side pauses and
    # come out up to ~1 % faster at N = 8 for that reason: profiles/r7b_*, r9e_*, r9f_*.)
    if rank == 0:
        sampler.wait_first_sample(5.0)
    fence()
    # (the decision to keep going is taken collectively: every rank must issue the same number
    # of steps, or the epochs of the fused exchange would drift apart between ranks)
    tw = time.perf_counter()
    while True:
        more = torch.tensor([1 if time.perf_counter() - tw < args.settle_s else 0], dtype=torch.int32, device=dev)
        if world > 1:
            dist.all_reduce(more, op=dist.ReduceOp.MIN)
        if int(more.item()) == 0:
            break
        for _ in range(50):
            step()
        fence()

    # Overlapped launches were validated at 1, 2 and 4 GPUs; should they ever misbehave on a box
    # (a peer that does not deliver ends in FLAGSTAT_CUDA_ETIMEOUT, not in a hang), all ranks fall
    # back together to serialised launches on a fresh exchange rather than lose the measurement.
    overlap_fallback = None
    if xchg is not None and overlap:
        ok = 1
        try:
            xchg.status()
        except Exception as exc:  # FlagstatCudaError
            ok = 0
            overlap_fallback = repr(exc)
        okt = torch.tensor([ok], dtype=torch.int32, device=dev)
        if world > 1:
            dist.all_reduce(okt, op=dist.ReduceOp.MIN)
        if int(okt.item()) == 0:
            overlap_fallback = overlap_fallback or "a peer reported a failed exchange"
            xchg.close()
            xchg = sharded.FusedExchange(device=dev, overlap=False)
            overlap = False
            deferred = False
            for _ in range(max(args.warmup, 3)):
                step()
            fence()

    # ---- device-resident timed region: exactly K steps ---------------------
    if xchg is not None:
        finish_steps()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = fs.lib().FLAGSTAT_cuda_launch_count()
    fence()
    t0 = time.perf_counter()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    if xchg is not None:
        finish_steps()  # the last step's counters (deferred collection): inside the timed region
    e1.record(stream)
    fence()
    t1 = time.perf_counter()
    launches = fs.lib().FLAGSTAT_cuda_launch_count() - launches0
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    result = counters.cpu().numpy().view(np.uint64).copy()
    if xchg is not None:
        xchg.status()

    # the same K steps with launches strictly serialised (no overlap of consecutive steps), for the record
    serial_ms = None
    serial_ok = None
    if xchg is not None and overlap:
        finish_steps()
        xchg.set_overlap(False)
        was_deferred, deferred = deferred, False  # one call = count + exchange + wait, strictly in order
        for _ in range(3):
            step()
        fence()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record(stream)
        for _ in range(args.steps):
            step()
        s1.record(stream)
        fence()
        t = torch.tensor([s0.elapsed_time(s1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        serial_ms = float(t.item()) / args.steps
        serial_ok = counters.cpu().numpy().view(np.uint64).tolist() == result.tolist()
        xchg.set_overlap(True)
        deferred = was_deferred

    # overlapped steps with the OTHER collection order, for the record (wait in the same launch if the
    # headline used deferred collection, and the other way round): same K steps, same stream
    immediate_ms = None
    deferred_ms = None
    if xchg is not None and overlap and world > 1:
        main_deferred = deferred
        finish_steps()
        deferred = not main_deferred
        for _ in range(3):
            step()
        finish_steps()
        fence()
        i0, i1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        i0.record(stream)
        for _ in range(args.steps):
            step()
        finish_steps()
        i1.record(stream)
        fence()
        t = torch.tensor([i0.elapsed_time(i1)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        other_ok = counters.cpu().numpy().view(np.uint64).tolist() == result.tolist()
        if deferred:
            deferred_ms = float(t.item()) / args.steps if other_ok else None
            immediate_ms = ms_total / args.steps
        else:
            immediate_ms = float(t.item()) / args.steps if other_ok else None
            deferred_ms = ms_total / args.steps
        deferred = main_deferred

    # the same K steps with the other exchange (kernel + separate NCCL all-reduce), for the record
    alt_ms = None
    if world > 1 and xchg is not None:
        for _ in range(3):
            step_nccl()
        fence()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record(stream)
        for _ in range(args.steps):
            step_nccl()
        a1.record(stream)
        fence()
        t = torch.tensor([a0.elapsed_time(a1)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        alt_ms = float(t.item()) / args.steps
        alt_ok = counters.cpu().numpy().view(np.uint64).tolist() == result.tolist()

    # ---- BASELINE configs[3]: 2^34 records range-sharded over the ranks (STRONG scaling) ----------
    # rank r holds shard_range(2^34, world, r) of the same generator (34.4 GB at N = 1); K overlapped
    # steps and K serialised ones, checked against the reference-made answer in the golden fixture
    strong = None
    if not args.no_strong:
        try:
            strong = strong_scaling_leg(args, fs, sharded, synth, xchg, overlap, dev, stream, world, rank, fence, dist, torch, np)
        except Exception as exc:  # the headline numbers must survive a failure of this leg
            strong = {"error": repr(exc)}

    # ---- the kernel alone (roofline): same stream, CUDA events -------------
    # Long enough (>= ~0.4 s) for nvidia-smi to see clocks and throttle reasons
    # under sustained load; the roofline number is the mean over this loop.
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    scratch = torch.zeros(32, dtype=torch.int64, device=dev)
    est_ms = max(ms_total / args.steps, 1e-3)
    kiters = max(args.steps, min(4000, int(400.0 / est_ms)))
    torch.cuda.synchronize(dev)
    tk0 = time.perf_counter()
    k0.record(stream)
    for _ in range(kiters):
        fs.flagstat_device(data, out=scratch, stream=stream)
    k1.record(stream)
    torch.cuda.synchronize(dev)
    tk1 = time.perf_counter()
    kernel_ms = k0.elapsed_time(k1) / kiters
    kmax = torch.tensor([kernel_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(kmax, op=dist.ReduceOp.MAX)
    kernel_ms_max = float(kmax.item())
    clocks = sampler.stop(tk0, tk1) if rank == 0 else None
    if clocks is not None:
        clocks["window"] = f"{kiters} back-to-back kernel launches ({(tk1 - tk0) * 1e3:.0f} ms) right after the timed steps"

    # ---- read-only HBM probe on the same bytes (LDG.128 + one XOR per 16 B, nothing else):
    # what a stream that only READS reaches on this device in this run; MEASURED_PEAKS.json
    # is a copy (read + write), which a read-only kernel can exceed
    read_probe_gbs = None
    try:
        import ctypes as C
        pm = C.c_float(0)
        nbytes16 = (2 * n) & ~15
        fs.check(fs.lib().FLAGSTAT_cuda_read_probe(data.data_ptr(), nbytes16, 3, C.byref(pm)), "read_probe")
        fs.check(fs.lib().FLAGSTAT_cuda_read_probe(data.data_ptr(), nbytes16, 200, C.byref(pm)), "read_probe")
        read_probe_gbs = nbytes16 / (pm.value * 1e-3) / 1e9
    except Exception as exc:
        read_probe_gbs = None
        print(f"bench.py: read probe failed: {exc!r}", file=sys.stderr)

    # ---- end to end through the public host-pointer API ---------------------
    host = torch.empty(n, dtype=torch.int16, pin_memory=True)
    host.copy_(data)
    torch.cuda.synchronize(dev)
    host_np = host.numpy().view(np.uint16)
    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(2):
        fs.flagstat_u64(host_np)
    fence()
    w0 = time.perf_counter()
    for _ in range(e2e_steps):
        f_e2e = fs.flagstat_u64(host_np)
    torch.cuda.synchronize(dev)
    w1 = time.perf_counter()
    e2e_s = torch.tensor([(w1 - w0) / e2e_steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_s = float(e2e_s.item())
    # PCIe probe: plain pinned cudaMemcpyAsync of the same buffer
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tmp = torch.empty_like(data)
    tmp.copy_(host, non_blocking=True)
    torch.cuda.synchronize(dev)
    p0.record(stream)
    tmp.copy_(host, non_blocking=True)
    p1.record(stream)
    torch.cuda.synchronize(dev)
    pcie_gbs = 2 * n / (p0.elapsed_time(p1) * 1e-3) / 1e9
    del tmp

    # ---- the same call on PAGEABLE memory (what numpy / malloc give the reference's callers):
    # T threads copy slices into pinned slots (run_pageable); N=1 only, rank 0
    pageable_info = None
    if world == 1:
        try:
            page_np = np.empty(n, np.uint16)
            page_np[:] = host_np
            fs.flagstat_u64(page_np)
            t0 = time.perf_counter()
            for _ in range(3):
                f_pg = fs.flagstat_u64(page_np)
            t1 = time.perf_counter()
            pg_s = (t1 - t0) / 3
            pageable_info = {
                "value": n / pg_s, "unit": UNIT, "gbs": 2 * n / pg_s / 1e9,
                "frac_of_pcie_probe": 2 * n / pg_s / 1e9 / pcie_gbs,
                "api": "FLAGSTAT_cuda_u64(pageable numpy array): threaded copy into pinned slots, "
                       "one DMA + launch per slice",
                "threads": int(os.environ.get("FLAGSTAT_CUDA_IO_THREADS", "0"))
                or max(2, min(24, len(os.sched_getaffinity(0)) * 3 // 4)),  # the library's default (io_threads())
                "same_counters": f_pg.tolist() == f_e2e.tolist(),
            }
            del page_np
        except Exception as exc:
            pageable_info = {"error": repr(exc)}

    # ---- streamed end to end (BASELINE configs[4]): 1,024,000-byte blocks from a pinned ring,
    # DMA of group k+1 overlapping the kernel of group k on separate streams; the submit loop
    # runs in C (FLAGSTAT_cuda_stream_selftime), data already in the pinned slots
    stream_info = None
    try:
        nblk_ring = 4 * 8
        with fs.BlockStream(local, fs.BLOCK_RECORDS, 4, mode=fs.BlockStream.DMA, coalesce=8) as bs:
            for i in range(nblk_ring):
                slot = bs.acquire()
                slot[:] = host_np[i * fs.BLOCK_RECORDS:(i + 1) * fs.BLOCK_RECORDS]
                bs.submit(fs.BLOCK_RECORDS)
            f_ring = bs.finish()
            bs.selftime(256)
            fence()
            laps = 40
            f_s, sec_s = bs.selftime(laps * nblk_ring)
        sec_t = torch.tensor([sec_s], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(sec_t, op=dist.ReduceOp.MAX)
        sec_s = float(sec_t.item())
        blocks = laps * nblk_ring
        stream_info = {
            "workload": f"{blocks} blocks x 1,024,000 B (512,000 records) per GPU from a pinned ring "
                        "(4 groups x 8 blocks), one DMA + one launch per group, 4 streams",
            "value": world * blocks * fs.BLOCK_RECORDS / sec_s, "unit": UNIT,
            "gbs_per_gpu": blocks * 1.024e-3 / sec_s, "pcie_h2d_probe_gbs": pcie_gbs,
            "frac_of_pcie_probe": blocks * 1.024e-3 / sec_s / pcie_gbs,
            "verified": f_s.tolist() == [laps * int(x) for x in f_ring],
            "note": "blocks are ALREADY in the pinned slots (FLAGSTAT_cuda_stream_selftime resubmits the ring): the "
                    "producer's fill is not in this number; with a host memcpy per block into the slot it was "
                    "9 - 11 GB/s on one thread (profiles/r1d_sweep.jsonl, r1h_sweep_tma_variants.jsonl: gbs_with_host_memcpy)",
        }
    except Exception as exc:  # the headline numbers must survive a failure of this extra leg
        stream_info = {"error": repr(exc)}

    # ---- N = 1 extras: configs[0] on the GPU and the FLAG-file readers ------------------
    inmemory_info = file_info = None
    if world == 1 and not args.no_extras:
        try:
            inmemory_info = inmemory_leg(fs, synth, torch, dev, load_peaks()[0])
        except Exception as exc:
            inmemory_info = {"error": repr(exc)}
        try:
            file_info = file_leg(fs, torch, dev, len(os.sched_getaffinity(0)), pcie_gbs)
        except Exception as exc:
            file_info = {"error": repr(exc)}

    # ---- verification: N x KAT-E, from the committed golden fixture ---------
    verified = None
    try:
        with open(os.path.join(ROOT, "tests", "golden", "flagstat_golden.json")) as fh:
            gold = json.load(fh)
            kat, kat16 = gold["kat_e"]["cuda_expected"], gold.get("kat_16g", {"spec": {"n": -1}})
        if n == HISEQX_N:
            # every rank's shard is one full period of the generator, so the global
            # answer is world x KAT-E and this rank's host-API answer is KAT-E
            verified = result.tolist() == [world * x for x in kat] and f_e2e.tolist() == kat
        else:
            # other shard sizes: every record counted exactly once, and -- for BASELINE
            # configs[3], 2^34 records in total -- the exact answer from the golden fixture
            # (20 periods of the generator + a 689,031,344-record prefix, made with the reference)
            total = world * n
            ok = int(result[9]) + int(result[25]) == total and int(f_e2e[9]) + int(f_e2e[25]) == n
            if total % HISEQX_N == 0:
                ok = ok and result.tolist() == [total // HISEQX_N * x for x in kat]
            elif total == kat16["spec"]["n"]:
                ok = ok and result.tolist() == kat16["cuda_expected"]
            verified = bool(ok)
    except Exception:
        verified = None

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    # ---- CPU baseline on this box's host cores (N=1 only) -------------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        sample = min(n, args.cpu_sample)
        r = cpu_reference_run(sample, 7, 2, threads)
        best = min(r["times"])
        cpu = {
            "value": sample / best, "unit": UNIT, "cores": r["threads"], "kind": r["kind"],
            "sample": f"first {sample} records of the workload, kernel {r['kernel']}, "
                      f"{r['threads']} pthreads over contiguous ranges, best of 7",
            "one_thread_value": sample / r["one_thread_s"], "cpu": cpu_model(),
            "verified": r["verified"],
        }
        if not args.no_extras:
            # scalar / avx2 / avx512 / avx512_improved3 / STORM_pospopcnt_u16 on the
            # 100 M-record U(0,4095) buffer, 1 thread and all cores; and the drop-in block loop
            try:
                cpu["inmemory_100m_kernel_set"] = cpu_baseline_set(threads)
            except Exception as exc:
                cpu["inmemory_100m_kernel_set"] = {"error": repr(exc)}
            try:
                cpu["dropin_block_loop"] = cpu_dropin_loop()
            except Exception as exc:
                cpu["dropin_block_loop"] = {"error": repr(exc)}

    peak, peak_src = load_peaks()
    step_ms = ms_total / args.steps
    value = world * n / (step_ms * 1e-3)
    # one call at a time (every launch waits for the previous one): what a single FLAGSTAT_cuda_device*
    # call costs; `value` pipelines consecutive calls (ms_per_step < roofline.kernel_ms is that overlap)
    value_serialised = (world * n / (serial_ms * 1e-3)) if serial_ms else None
    achieved = 2.0 * n / (kernel_ms * 1e-3) / 1e9
    traffic = load_traffic()
    kernel_name = fs.lib().FLAGSTAT_cuda_kernel_name(0).decode()
    if traffic and "fsb200::" + traffic.get("kernel", "") != kernel_name:
        traffic = None  # the committed ncu capture is of another kernel: no traffic claim
    line = {
        "metric": METRIC, "value": value, "value_serialised": value_serialised, "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": step_ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u16", "data": "synthetic",
        "config": workload_config(world, n),
        "roofline": {
            "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": achieved / peak, "peak_source": peak_src,
            "kernel": kernel_name,
            "kernel_ms": kernel_ms, "kernel_ms_slowest_rank": kernel_ms_max,
            "read_only_probe_gbs": read_probe_gbs,
            "frac_of_read_only_probe": (achieved / read_probe_gbs) if read_probe_gbs else None,
            "algorithmic_bytes_per_launch": 2 * n,
            # the same bytes over the timed step (launch gaps, kernel tail and exchange included;
            # with overlapped steps this exceeds `achieved`, which times launches one by one)
            "step_gbs_per_gpu": 2.0 * n / (step_ms * 1e-3) / 1e9,
            "traffic": (traffic or {}).get("dram_bytes_per_launch"),
            "traffic_source": (traffic or {}).get("source"),
        },
        "cpu_baseline": cpu,
        "e2e": {
            "value": world * n / e2e_s, "unit": UNIT, "h2d_bytes_per_step": 2 * n,
            "d2h_bytes_per_step": 256, "ms_per_step": e2e_s * 1e3,
            "api": "FLAGSTAT_cuda_u64(pinned host pointer) via libflagstats_b200.flagstat_u64",
            "achieved_gbs_per_gpu": 2 * n / e2e_s / 1e9, "pcie_h2d_probe_gbs": pcie_gbs,
            "frac_of_pcie_probe": (2 * n / e2e_s / 1e9) / pcie_gbs,
        },
        "e2e_pageable": pageable_info,
        "stream_e2e": stream_info,
        "strong_2p34": strong,
        "inmemory_100m": inmemory_info,
        "file_e2e": file_info,
        "gpu_launches": int(launches),
        "exchange": ("fused: counters exchanged by the counting kernel itself through peer-mapped "
                     "memory (FLAGSTAT_cuda_device_allreduce), 1 launch/step"
                     + ("; consecutive steps overlap (programmatic dependent launch: the next step "
                        "streams its shard while this step's last CTA exchanges counters)" if overlap else "")
                     + ("; deferred collection: a step pushes its totals, the next step's launch (the last "
                        "one: FLAGSTAT_cuda_xchg_collect, inside the timed region) waits for the peers' and "
                        "writes the counters" if deferred and world > 1 else "")
                     if xchg is not None else "kernel + memset + NCCL all-reduce of 32 x u64"),
        "overlap_fallback": overlap_fallback,
        "ms_per_step_serialised_launches": serial_ms,
        "ms_per_step_overlapped_wait_in_launch": immediate_ms,
        "ms_per_step_overlapped_deferred_collection": deferred_ms,
        "serialised_same_result": serial_ok,
        "ms_per_step_with_nccl_allreduce": alt_ms,
        "nccl_path_same_result": (alt_ok if alt_ms is not None else None),
        "clocks": clocks,
        "verified": verified,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--records", type=int, default=HISEQX_N, help="records per GPU")
    ap.add_argument("--cpu-sample", type=int, default=200_000_000,
                    help="records of the workload the CPU baseline is timed on")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--settle-s", type=float, default=0.3,
                    help="extra warm-up under load before the timed steps (seconds)")
    ap.add_argument("--no-overlap", action="store_true",
                    help="launch consecutive fused steps strictly serialised (no programmatic dependent launch)")
    ap.add_argument("--deferred", action="store_true",
                    help="headline steps with deferred collection (push only; the next launch collects) instead of "
                         "waiting for the peers' totals in the same launch; the other order is always measured too")
    ap.add_argument("--no-deferred", action="store_true", help="(the default since session r7b; accepted and ignored)")
    ap.add_argument("--no-strong", action="store_true", help="skip the 2^34-record strong-scaling leg (configs[3])")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the N=1 extra legs (100 M in-memory config, FLAG files, drop-in block loop, CPU kernel set)")
    ap.add_argument("--exchange", choices=["fused", "nccl"], default="fused",
                    help="how the 32 counters are summed across ranks")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)
    return ours(args)


if __name__ == "__main__":
    sys.exit(main())
