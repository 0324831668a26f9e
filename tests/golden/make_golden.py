#!/usr/bin/env python3
"""Generate tests/golden/flagstat_golden.json by RUNNING THE UNMODIFIED
REFERENCE (oracle/_ref, built from /root/reference by oracle/Makefile).

Run in the dev container:   python tests/golden/make_golden.py
The GPU box has no /root/reference; its tests regenerate every input from the
generator spec stored with each case (oracle/flagstat_oracle.c generators or
the closed forms below) and compare against the numbers stored here.

Every expected vector in the output comes from reference code:
  scalar    -> FLAGSTAT_scalar                 (libflagstats.h:170-176)
  dispatch  -> FLAGSTATS_u16                   (libflagstats.h:3024-3070)
  simd      -> agreed CORE19 + slot 9 of FLAGSTAT_sse4/avx2/avx512 (those the
               host can run), asserted identical before being written
  pospopcnt -> STORM_pospopcnt_u16             (libalgebra.h:3496-3551)
  samtools  -> flagstat_loop into bam_flagstat_t (benchmark/flagstats.cpp:43-71), the
               macro itself compiled into oracle/_ref; 13 x [pass, fail]
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

CORE19 = list(O.CORE19)
CORE20 = list(O.CORE20)
SIMD = ("sse4", "avx2", "avx512")

SWEEP_LENGTHS = [0, 1, 2, 7, 8, 9, 15, 16, 17, 31, 33, 63, 64, 65, 127, 128, 129, 255, 256, 257,
                 511, 512, 513, 1000, 1023, 1024, 1025, 4095, 4096, 4097, 12345, 65535, 65536,
                 65537, 100000, 512000, 1000003, 16777216 + 777]


def mt19937_inmemory(n):
    """benchmark/inmemory.cpp:108-116: mt19937 seeded 0, libstdc++
    uniform_int_distribution<uint16_t>(0, 4095) == raw >> 20."""
    bg = np.random.MT19937()
    bg._legacy_seeding(0)
    return (bg.random_raw(n) >> 20).astype(np.uint16)


def make_input(spec):
    g = spec["gen"]
    if g == "arange":
        return (np.arange(spec["n"], dtype=np.uint32) & 0xFFFF).astype(np.uint16)
    if g == "mt19937_inmemory":
        return mt19937_inmemory(spec["n"])
    if g == "uniform":
        return O.synth_uniform(spec.get("start", 0), spec["n"], spec["seed"], spec["mask"])
    if g == "hiseqx":
        return O.synth_hiseqx(spec.get("start", 0), spec["n"], spec.get("seed", 0),
                              spec.get("qcfail_ppm", 0))
    if g == "const":
        return np.full(spec["n"], spec["value"], np.uint16)
    raise ValueError(g)


def reference_answers(a, valid_sam):
    lib = O.reference()
    assert lib is not None, "build oracle/_ref first (make -C oracle)"
    scalar = O.ref_flagstat("scalar", a).astype(np.uint64)
    out = {"scalar": scalar.tolist()}
    simd = None
    for k in SIMD:
        if not lib.ref_kernel_runnable(k.encode()):
            continue
        r = O.ref_flagstat(k, a).astype(np.uint64)
        if valid_sam:  # OR-based kernels leak bits >= 12
            assert (r[CORE19] == scalar[CORE19]).all(), (k, a.size)
            assert int(r[9]) == a.size - int(scalar[25]), (k, a.size)
            simd = r if simd is None else simd
            assert (r[CORE20] == simd[CORE20]).all()
    # improved3 stays exact with bits >= 12 set and equals scalar on all 32 slots
    if lib.ref_kernel_runnable(b"avx512_improved3"):
        r3 = O.ref_flagstat("avx512_improved3", a).astype(np.uint64)
        assert (r3 == scalar).all(), a.size
    expect = scalar.copy()
    expect[9] = np.uint64(a.size - int(scalar[25]))  # :429 / :1212 / :1843
    if simd is not None:
        assert (expect[CORE20] == simd[CORE20]).all()
    out["cuda_expected"] = expect.tolist()  # CORE19 from scalar + slot 9 SIMD convention
    if valid_sam:
        out["dispatch"] = O.ref_flagstats_u16(a).astype(np.uint64).tolist()
        out["dispatch_kernel"] = O.ref_dispatch_name(a.size)
    out["pospopcnt"] = O.ref_pospopcnt(a).astype(np.uint64).tolist()
    st = O.ref_samtools_loop(a)
    assert st is not None, "oracle/_ref predates ref_samtools_loop: make -C oracle"
    # ten of the eleven FLAG-derived fields follow from the counters (any input: the loop
    # tests bits 0..11 only, like the scalar rule)
    implied = O.samtools_from_counters(expect)
    implied[2] = st[2]
    assert (implied == st).all(), a.size
    out["samtools"] = st.tolist()
    return out


def core20_only(f):
    """Keep CORE19 + slot 9; the SIMD kernels also leave length-dependent raw-bit
    counts in slots 0,1,3,4,5 (body only, tail excluded) which are an accident of
    not masking, not a contract."""
    out = np.zeros(32, np.uint64)
    out[CORE20] = np.asarray(f, np.uint64)[CORE20]
    return out


def main():
    cases = []

    def add(name, spec, valid_sam=True):
        a = make_input(spec)
        c = {"name": name, "spec": spec, "valid_sam": valid_sam}
        c.update(reference_answers(a, valid_sam))
        cases.append(c)
        return a

    add("KAT-A all 12-bit values once", {"gen": "arange", "n": 4096})
    add("KAT-B all 16-bit values once", {"gen": "arange", "n": 65536}, valid_sam=False)
    add("KAT-D inmemory.cpp default input", {"gen": "mt19937_inmemory", "n": 102400})
    add("inmemory.cpp input, 1,000,000", {"gen": "mt19937_inmemory", "n": 1000000})

    for v in (0, 1, 3, 4, 0x200, 0x203, 0x100, 0x800, 0x900, 0xFFF, 0x4FF, 99, 147, 2113):
        add(f"const 0x{v:03x}", {"gen": "const", "n": 5000, "value": v})

    for n in SWEEP_LENGTHS:
        add(f"uniform12 n={n}", {"gen": "uniform", "n": n, "seed": 0, "mask": 0x0FFF})
    for n in (1, 255, 1024, 4097, 100000, 1000003):
        add(f"uniform16 n={n}", {"gen": "uniform", "n": n, "seed": 7, "mask": 0xFFFF},
            valid_sam=False)
    for n in (1000, 65536, 1000003):
        add(f"hiseqx prefix n={n}", {"gen": "hiseqx", "n": n})
        add(f"hiseqx qcfail n={n}", {"gen": "hiseqx", "n": n, "seed": 3, "qcfail_ppm": 20000})
    add("uniform12 shard start=10^9", {"gen": "uniform", "n": 70001, "seed": 5, "mask": 0x0FFF,
                                       "start": 1_000_000_007})

    # KAT-C: accumulate A then its first 2048 values into the same flags
    a = make_input({"gen": "arange", "n": 4096})
    f = O.ref_flagstat("scalar", a)
    f = O.ref_flagstat("scalar", a[:2048], f)
    kat_c = f.astype(np.uint64)
    kat_c[9] = np.uint64(4096 + 2048 - int(kat_c[25]))
    if O.reference().ref_kernel_runnable(b"avx512"):
        g = O.ref_flagstat("avx512", a)
        g = O.ref_flagstat("avx512", a[:2048], g).astype(np.uint64)
        assert (g[CORE20] == kat_c[CORE20]).all()

    # KAT-E: full HiSeqX-shaped column, 824,541,892 records, through the reference
    N = O.HISEQX_N
    print("generating KAT-E input (1.65 GB) ...", flush=True)
    e = make_input({"gen": "hiseqx", "n": N})
    vals, cnts = np.unique(e, return_counts=True)
    kernel = O.best_reference_kernel()
    fe, sec = O.ref_flagstat_mt(kernel, e, os.cpu_count() or 1)
    fs, _ = O.ref_flagstat_mt("scalar", e, os.cpu_count() or 1)
    assert (fe[CORE19] == fs[CORE19]).all()
    assert (O.numpy_flagstat(e)[CORE20] == fe[CORE20]).all()
    readme = {  # README.md:179-191 (samtools flagstat NA12878D_HiSeqX_R12_GRCh37.bam)
        "total": 824541892, "secondary": 0, "supplementary": 5393628, "duplicates": 0,
        "mapped": 805383403, "paired": 819148264, "read1": 409574132, "read2": 409574132,
        "properly_paired": 781085884, "both_mapped": 797950890, "singletons": 2038885,
    }
    assert int(fe[9]) + int(fe[25]) == readme["total"]
    assert int(fe[8]) == readme["secondary"] and int(fe[11]) == readme["supplementary"]
    assert int(fe[10]) == readme["duplicates"]
    assert readme["total"] - int(fe[2]) == readme["mapped"]
    assert int(fe[6]) == readme["read1"] and int(fe[7]) == readme["read2"]
    assert int(fe[6]) + int(fe[7]) == readme["paired"]
    assert int(fe[12]) == readme["properly_paired"] and int(fe[14]) == readme["both_mapped"]
    assert int(fe[13]) == readme["singletons"]
    assert (fe[16:] == 0).all()
    # pospopcnt over the full column, in shards of 2^30 (uint32 counters)
    pp = np.zeros(16, np.uint64)
    for lo in range(0, N, 1 << 30):
        pp += O.ref_pospopcnt(e[lo:lo + (1 << 30)]).astype(np.uint64)
    print("flagstat_loop over KAT-E ...", flush=True)
    st_e = O.ref_samtools_loop(e)
    assert int(st_e[2, 0]) == readme["paired"] and int(st_e[2, 1]) == 0
    implied = O.samtools_from_counters(core20_only(fe))
    implied[2] = st_e[2]
    assert (implied == st_e).all()
    report_e = O.samtools_report(st_e)
    # README.md:179-189 verbatim (the two diffchr lines are not FLAG-derivable)
    assert report_e == (
        "824541892 + 0 in total (QC-passed reads + QC-failed reads)\n0 + 0 secondary\n"
        "5393628 + 0 supplementary\n0 + 0 duplicates\n805383403 + 0 mapped (97.68% : N/A)\n"
        "819148264 + 0 paired in sequencing\n409574132 + 0 read1\n409574132 + 0 read2\n"
        "781085884 + 0 properly paired (95.35% : N/A)\n797950890 + 0 with itself and mate mapped\n"
        "2038885 + 0 singletons (0.25% : N/A)\n")
    kat_e = {
        "name": "KAT-E HiSeqX-shaped, README.md:179-191",
        "samtools": st_e.tolist(), "samtools_report": report_e,
        "spec": {"gen": "hiseqx", "n": N},
        "category_values": vals.tolist(), "category_counts": cnts.tolist(),
        "reference_kernel": kernel,
        "cuda_expected": core20_only(fe).tolist(), "reference_avx512_all32": fe.tolist(),
        "pospopcnt": pp.tolist(), "readme": readme,
        # shard-level answers so multi-GPU tests can check each range
        "shards8": [],
    }
    for g in range(8):
        lo = (g * N // 8) & ~7
        hi = ((g + 1) * N // 8) & ~7 if g < 7 else N
        fg, _ = O.ref_flagstat_mt(kernel, e[lo:hi], os.cpu_count() or 1)
        kat_e["shards8"].append({"start": lo, "n": hi - lo,
                                 "cuda_expected": core20_only(fg).tolist()})
    assert (np.sum([np.array(s["cuda_expected"], np.uint64) for s in kat_e["shards8"]], axis=0)[CORE20]
            == fe[CORE20]).all()

    # BASELINE configs[3]: 2^34 records of the same periodic generator = 20 periods + a prefix
    n16 = 1 << 34
    q16, r16 = divmod(n16, N)
    fp, _ = O.ref_flagstat_mt(kernel, e[:r16], os.cpu_count() or 1)
    fps, _ = O.ref_flagstat_mt("scalar", e[:r16], os.cpu_count() or 1)
    assert (fp[CORE19] == fps[CORE19]).all()
    kat_16g = {
        "name": "BASELINE configs[3]: 2^34 HiSeqX-shaped records (20 periods + 689,031,344)",
        "spec": {"gen": "hiseqx", "n": n16}, "periods": q16, "prefix": r16,
        "prefix_cuda_expected": core20_only(fp).tolist(),
        "cuda_expected": (np.uint64(q16) * core20_only(fe) + core20_only(fp)).tolist(),
        "n_pair_all": int(q16 * int(st_e[2, 0]) + int(O.ref_samtools_loop(e[:r16])[2, 0])),
    }

    doc = {
        "generated_by": "tests/golden/make_golden.py",
        "reference": "mklarqvist/libflagstats @93f68238 (libalgebra @bff182e8), unmodified, "
                     "compiled by oracle/Makefile; host kernel set: " + ",".join(O.ref_kernels()),
        "hiseqx": {"N": N, "M": int(O.oracle().oracle_hiseqx_m())},
        "core19": CORE19, "core20": CORE20,
        "cases": cases,
        "kat_c": {"cuda_expected": kat_c.tolist()},
        "kat_e": kat_e,
        "kat_16g": kat_16g,
    }
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "flagstat_golden.json")
    with open(path, "w") as fh:
        json.dump(doc, fh, indent=None, separators=(",", ":"))
        fh.write("\n")
    print("wrote", path, os.path.getsize(path), "bytes;", len(cases), "cases")


if __name__ == "__main__":
    main()
