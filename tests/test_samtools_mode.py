"""The reference benchmark's samtools-style caller: bam_flagstat_t filled by flagstat_loop
and printed as samtools' report (benchmark/flagstats.cpp:43-71, 73-78, 577-588).

CPU part: the oracle restatement against the reference's OWN macro (cut out of
flagstats.cpp into oracle/_ref at build time) and against the golden vectors made
with it; the host-only C entries (from_counters, report).
GPU part (-m gpu): FLAGSTAT_cuda_samtools* through the C ABI against the oracle.
"""
import ctypes as C

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from oracle import oracle as O
from tests.helpers import CORE20, make_input, offset_copy

ref = O.reference()
has_ref_loop = ref is not None and hasattr(ref, "ref_samtools_loop")
needs_ref = pytest.mark.skipif(not has_ref_loop, reason="oracle/_ref without ref_samtools_loop")


def expected_counters(a):
    """What FLAGSTAT_cuda_samtools_u64 must leave in zeroed flags[32]."""
    f = O.flagstat_simd(a)
    s = O.samtools_loop(a)
    f[0], f[16] = np.uint64(s[2, 0]), np.uint64(s[2, 1])
    return f, s


# ----------------------------------------------------------------------------- CPU
@needs_ref
@settings(max_examples=60, deadline=None)
@given(n=st.integers(0, 60000), seed=st.integers(0, 2**32),
       mask=st.sampled_from([0x0FFF, 0xFFFF, 0x03FF, 0x0DFF, 0x0EFF]))
def test_oracle_loop_is_the_references_macro(n, seed, mask):
    a = O.synth_uniform(0, n, seed, mask)
    assert O.samtools_loop(a).tolist() == O.ref_samtools_loop(a).tolist()


@needs_ref
def test_oracle_percent_is_the_references():
    rng = np.random.default_rng(5)
    pairs = [(0, 0), (0, 7), (7, 7), (1, 3), (2, 3), (805383403, 824541892), (2038885, 819148264),
             (1, 200), (1, 201), (5, 1000), (15, 1000), (25, 1000), (2**40, 2**41 + 1)]
    pairs += [(int(a), int(a + b)) for a, b in rng.integers(0, 10**9, (300, 2))]
    for n, t in pairs:
        assert O.samtools_percent(n, t) == O.ref_samtools_percent(n, t), (n, t)


def test_oracle_loop_against_golden(golden):
    for c in golden["cases"]:
        if c["spec"]["n"] > 200_000:
            continue
        a = make_input(c["spec"])
        assert O.samtools_loop(a).tolist() == c["samtools"], c["name"]


def test_loop_accumulates():
    a = O.synth_uniform(0, 5000, 3, 0x0FFF)
    s = O.samtools_loop(a)
    s = O.samtools_loop(a[:777], s)
    assert s.tolist() == (O.samtools_loop(a) + O.samtools_loop(a[:777])).tolist()


def test_ten_fields_follow_from_the_counters_and_n_pair_all_does_not():
    a = O.synth_uniform(0, 40000, 11, 0x0FFF)
    f, s = expected_counters(a)
    assert O.samtools_from_counters(f).tolist() == s.tolist()
    # READ1 + READ2 (python/libflagstats.pyx:35) is NOT n_pair_all on arbitrary FLAG words
    assert int(f[6]) + int(f[7]) != int(s[2, 0])
    # ... but is on well-formed pairs (exactly one of READ1 / READ2 per paired record)
    h = O.synth_hiseqx(0, 40000, 0, 20000)
    fh, sh = expected_counters(h)
    assert int(fh[6]) + int(fh[7]) == int(sh[2, 0]) and int(fh[22]) + int(fh[23]) == int(sh[2, 1])


def test_c_entries_that_need_no_device(golden):
    """from_counters and report are host code: exercised here against the oracle."""
    import libflagstats_b200 as fs
    for spec in ({"gen": "uniform", "n": 30000, "seed": 2, "mask": 0xFFFF},
                 {"gen": "hiseqx", "n": 50000, "seed": 3, "qcfail_ppm": 20000},
                 {"gen": "const", "n": 10, "value": 0x100}, {"gen": "const", "n": 0, "value": 0}):
        a = make_input(spec)
        f, s = expected_counters(a)
        got = fs.samtools_stats_from_counters(f)
        assert got.tolist() == s.tolist()
        assert fs.samtools_text(got) == O.samtools_report(s)
    e = golden["kat_e"]
    s = np.array(e["samtools"], np.int64)
    f = np.array(e["cuda_expected"], np.uint64)
    f[0] = np.uint64(e["readme"]["paired"])
    assert fs.samtools_stats_from_counters(f).tolist() == s.tolist()
    assert fs.samtools_text(s) == e["samtools_report"]
    # from_counters ADDS
    acc = s.copy()
    fs.check(fs.lib().FLAGSTAT_cuda_samtools_from_counters(f.ctypes.data_as(C.POINTER(C.c_uint64)),
                                                           acc.ctypes.data), "from_counters")
    assert acc.tolist() == (2 * s).tolist()
    # snprintf convention
    buf = C.create_string_buffer(16)
    n = fs.lib().FLAGSTAT_cuda_samtools_report(s.ctypes.data, buf, 16)
    assert n == len(e["samtools_report"]) and buf.value.decode() == e["samtools_report"][:15]
    assert fs.lib().FLAGSTAT_cuda_samtools_report(None, buf, 16) == -2


def test_no_cpu_fallback_in_samtools_entries():
    import libflagstats_b200 as fs
    if fs.available() > 0:
        pytest.skip("a device is present")
    a = np.zeros(10, np.uint16)
    with pytest.raises(fs.FlagstatCudaError) as ei:
        fs.samtools_stats(a)
    assert ei.value.code == -1
    with pytest.raises(fs.FlagstatCudaError):
        fs.flagstat_samtools_u64(a)


# ----------------------------------------------------------------------------- GPU
@pytest.mark.gpu
def test_every_record_value_in_both_register_halves_samtools(cuda_lib):
    """All 65,536 FLAG words at even and odd record positions, alone and mixed."""
    fs = cuda_lib
    w = np.arange(65536, dtype=np.uint32).astype(np.uint16)
    for a in (w, np.concatenate([w[:1], w]), np.repeat(w, 3), np.tile(w, 40)[::-1].copy()):
        f, s = expected_counters(a)
        assert fs.flagstat_samtools_u64(a).tolist() == f.tolist()
        assert fs.samtools_stats(a).tolist() == s.tolist()
    # single-class columns: each record value 4096 times (a whole number of warp batches
    # takes the no-SECONDARY / SECONDARY / QC-fail paths separately)
    for v in (0x001, 0x011, 0x005, 0x045, 0x083, 0x101, 0x801, 0x901, 0x201, 0x205, 0x301, 0xB01,
              0x000, 0x010, 0x210, 0xFFF, 0x6FF, 0x0FB, 0xF01B):
        a = np.full(300_000, v, np.uint16)
        f, s = expected_counters(a)
        assert fs.flagstat_samtools_u64(a).tolist() == f.tolist(), hex(v)


@pytest.mark.gpu
def test_samtools_golden_cases_host_and_device(cuda_lib, golden):
    import torch
    fs = cuda_lib
    for c in golden["cases"]:
        a = make_input(c["spec"])
        want = np.array(c["cuda_expected"], np.uint64)
        st_ = np.array(c["samtools"], np.int64)
        want[0], want[16] = np.uint64(st_[2, 0]), np.uint64(st_[2, 1])
        assert fs.flagstat_samtools_u64(a).tolist() == want.tolist(), c["name"]
        d = torch.from_numpy(a.view(np.int16)).cuda()
        assert fs.samtools_stats(d).tolist() == st_.tolist(), c["name"]


@pytest.mark.gpu
def test_samtools_random_lengths_offsets_and_accumulate(cuda_lib):
    import torch
    fs = cuda_lib
    rng = np.random.default_rng(17)
    acc = np.zeros(32, np.uint64)
    acc_want = np.zeros(32, np.uint64)
    for i in range(40):
        n = int(rng.choice([0, 1, 7, 9, 511, 4097, 70001, 1_000_003, 3_000_017]))
        off = int(rng.integers(0, 8))
        mask = int(rng.choice([0x0FFF, 0xFFFF, 0x0CFF, 0x0DFF]))
        a = offset_copy(O.synth_uniform(i * 1000, n, i, mask), off)
        f, s = expected_counters(a)
        fs.flagstat_samtools_u64(a, acc)
        acc_want += f
        assert acc.tolist() == acc_want.tolist(), (i, n, off)
        dev = torch.from_numpy(np.concatenate([np.zeros(off, np.uint16), a]).view(np.int16)).cuda()[off:]
        out = torch.zeros(32, dtype=torch.int64, device="cuda")
        fs.check(fs.lib().FLAGSTAT_cuda_samtools_device(dev.data_ptr(), n, out.data_ptr(),
                                                        torch.cuda.current_stream().cuda_stream),
                 "FLAGSTAT_cuda_samtools_device")
        torch.cuda.synchronize()
        assert out.cpu().numpy().view(np.uint64).tolist() == f.tolist(), (i, n, off)


@pytest.mark.gpu
def test_kat_e_samtools_report_is_the_readme(cuda_lib, golden):
    """BASELINE configs[1]: the full HiSeqX-shaped column gives README.md:179-189 verbatim,
    'paired in sequencing' included."""
    fs = cuda_lib
    from libflagstats_b200 import synth
    e = golden["kat_e"]
    d = synth.hiseqx_device(O.HISEQX_N)
    s = fs.samtools_stats(d)
    assert s.tolist() == e["samtools"]
    assert fs.samtools_text(s) == e["samtools_report"]
    f = fs.flagstat_samtools_u64(d)
    assert f[CORE20].tolist() == np.array(e["cuda_expected"], np.uint64)[CORE20].tolist()
    assert int(f[0]) == e["readme"]["paired"] and int(f[16]) == 0
    # the plain contract still leaves slots 0 / 16 alone
    g = fs.flagstat_u64(d)
    assert int(g[0]) == 0 and int(g[16]) == 0
    # a QC-fail mix of the same shape: pass + fail == clean, n_pair_all included
    q = synth.hiseqx_device(60_000_000, 0, 9, 30000)
    sq = fs.samtools_stats(q)
    sc = fs.samtools_stats(d[:60_000_000])
    assert (sq[:, 0] + sq[:, 1]).tolist() == (sc[:, 0] + sc[:, 1]).tolist() and int(sq[0, 1]) > 0


@pytest.mark.gpu
def test_samtools_files_raw_and_lz4(cuda_lib, tmp_path):
    """The reference's 'samtools' readers of its FLAG files (flagstats.cpp:496-519, 547-590)."""
    fs = cuda_lib
    from libflagstats_b200 import blockfile
    a = O.synth_hiseqx(0, 3 * 512_000 + 12345, 4, 15000)
    f, _ = expected_counters(a)
    raw = tmp_path / "flags.bin"
    a.tofile(raw)
    got, n = blockfile.flagstat_file(str(raw), blockfile.RAW, samtools=True)
    assert n == a.size and got.tolist() == f.tolist()
    blob = O.write_lz4_container(a)
    got, n = blockfile.flagstat_container(blob, blockfile.LZ4, samtools=True)
    assert n == a.size and got.tolist() == f.tolist()
    lz = tmp_path / "flags.lz4"
    lz.write_bytes(blob)
    got, n = blockfile.flagstat_file(str(lz), blockfile.LZ4, samtools=True)
    assert n == a.size and got.tolist() == f.tolist()
    # without the flag the plain contract holds (slots 0 / 16 untouched)
    got, _ = blockfile.flagstat_container(blob, blockfile.LZ4)
    assert int(got[0]) == 0 and got[CORE20].tolist() == f[CORE20].tolist()
