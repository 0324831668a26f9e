"""The drop-in boundary, exercised through the REFERENCE's own entry points.

oracle/_ref/dropin_check and oracle/_ref/pyflagstats*.so are built by
integration/build_dropin.sh from the reference's libflagstats.h with
integration/libflagstats_h_cuda.patch applied and from its UNCHANGED
python/libflagstats.pyx.  Without a GPU the patched dispatcher must fall through
to the reference's CPU kernels; with one it must route long blocks to
FLAGSTAT_cuda -- and give the same counters either way."""
import importlib.util
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")
CHECK = os.path.join(REF_DIR, "dropin_check")


def _pyflagstats():
    for f in os.listdir(REF_DIR) if os.path.isdir(REF_DIR) else []:
        if f.startswith("pyflagstats") and f.endswith(".so"):
            spec = importlib.util.spec_from_file_location("pyflagstats", os.path.join(REF_DIR, f))
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            return mod
    return None


def _have_gpu():
    import torch
    return torch.cuda.is_available()


def _check_dict(d, a):
    want = O.flagstat_simd(a)
    names = ["FPAIRED", "FPROPER_PAIR", "FUNMAP", "FMUNMAP", "FREVERSE", "FMREVERSE", "FREAD1",
             "FREAD2", "FSECONDARY", "FQCFAIL", "FDUP", "FSUPPLEMENTARY", "n_pair_good", "n_sgltn",
             "n_pair_map"]
    assert d["n_values"] == a.size
    for j in (2, 6, 7, 8, 10, 11, 12, 13, 14):
        assert int(d["passed"][names[j]]) == int(want[j]), names[j]
        assert int(d["failed"][names[j]]) == int(want[16 + j]), names[j]
    assert int(d["failed"]["FQCFAIL"]) == int(want[25])
    assert int(d["passed"]["FQCFAIL"]) == int(want[9])  # n >= 256: SIMD / CUDA convention
    assert int(d["passed"]["mapped"]) == a.size - int(want[2]) - int(want[18])


@pytest.mark.skipif(not os.path.exists(CHECK), reason="integration/build_dropin.sh not run")
def test_patched_reference_dispatch_without_gpu():
    if _have_gpu():
        pytest.skip("GPU present: covered by the gpu-marked test")
    r = subprocess.run([CHECK], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "devices=0" in r.stdout and "OK (cuda selected for 0 calls)" in r.stdout
    assert "FLAGSTAT_avx512" in r.stdout or "FLAGSTAT_avx2" in r.stdout or "FLAGSTAT_sse4" in r.stdout


@pytest.mark.skipif(not os.path.exists(CHECK), reason="integration/build_dropin.sh not run")
def test_unchanged_pyx_against_patched_header_without_gpu():
    if _have_gpu():
        pytest.skip("GPU present: covered by the gpu-marked test")
    mod = _pyflagstats()
    assert mod is not None
    a = O.synth_uniform(0, 1_000_003, 3, 0x0FFF)
    _check_dict(mod.flagstats(a), a)
    with pytest.raises(ValueError):
        mod.flagstats([1, 2, 3])


@pytest.mark.gpu
def test_patched_reference_dispatch_selects_flagstat_cuda():
    assert os.path.exists(CHECK), "oracle/_ref/dropin_check missing (integration/build_dropin.sh)"
    r = subprocess.run([CHECK], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    # default FLAGSTAT_cuda_min_len = 1,048,576 (the measured pageable-memory crossover): long columns go to
    # the device, the reference's own 512,000-record block stays on the CPU kernels ...
    assert "n=5300123  -> FLAGSTAT_cuda" in r.stdout, r.stdout
    assert "n=512000   -> FLAGSTAT_avx" in r.stdout or "n=512000   -> FLAGSTAT_sse4" in r.stdout, r.stdout
    assert "n=65536    -> FLAGSTAT_avx" in r.stdout or "n=65536    -> FLAGSTAT_sse4" in r.stdout, r.stdout
    # ... and with FLAGSTAT_cuda_set_min_len(4096) every block of the loop (11 blocks) goes to the device:
    # dropin_check fails ("only N blocks went to FLAGSTAT_cuda") otherwise
    assert "MISMATCH" not in r.stdout and "FAIL" not in r.stdout, r.stdout
    assert "OK (cuda selected for" in r.stdout
    selected = int(r.stdout.rsplit("cuda selected for", 1)[1].split()[0])
    assert selected >= 11 + 2, r.stdout


@pytest.mark.gpu
def test_unchanged_pyx_runs_on_the_gpu():
    mod = _pyflagstats()
    assert mod is not None, "oracle/_ref/pyflagstats*.so missing (integration/build_dropin.sh)"
    import libflagstats_b200 as fs
    before = fs.lib().FLAGSTAT_cuda_launch_count()
    # both lengths must reach the device (pyflagstats links the same libflagstats_cuda.so this process
    # has loaded): the default threshold would leave the first one to FLAGSTAT_avx512, whose dict differs
    # from ours in the slots that kernel fills with raw-bit counts (slots 0,1,3,4,5)
    prev_min = fs.lib().FLAGSTAT_cuda_min_len()
    fs.lib().FLAGSTAT_cuda_set_min_len(4096)
    try:
        _pyx_checks(mod, fs)
    finally:
        fs.lib().FLAGSTAT_cuda_set_min_len(prev_min)
    assert fs.lib().FLAGSTAT_cuda_launch_count() > before  # same process, same .so: CUDA really ran


def _pyx_checks(mod, fs):
    for n in (1_000_003, 100_000_000):
        a = O.synth_uniform(0, n, 3, 0x0FFF)
        d = mod.flagstats(a)  # the reference's own Python entry point
        ours = fs.flagstats(a)
        if n <= 1_000_003:
            _check_dict(d, a)
        assert {k: int(v) for k, v in d["passed"].items()} == {k: int(v) for k, v in ours["passed"].items()}
        assert {k: int(v) for k, v in d["failed"].items()} == {k: int(v) for k, v in ours["failed"].items()}


SAMCALL = os.path.join(REF_DIR, "samtools_caller")


@pytest.mark.skipif(not os.path.exists(SAMCALL), reason="integration/build_dropin.sh not run")
def test_plain_c_samtools_caller_has_no_cpu_fallback(tmp_path):
    """integration/samtools_caller.c (C99, only include/flagstats_cuda.h): without a device the
    reference's block loop around FLAGSTAT_cuda_samtools must fail, not count on the CPU."""
    if _have_gpu():
        pytest.skip("GPU present: covered by the gpu-marked test")
    p = tmp_path / "flags.bin"
    O.synth_uniform(0, 10_000, 1, 0x0FFF).tofile(p)
    r = subprocess.run([SAMCALL, str(p)], capture_output=True, text=True, timeout=60)
    assert r.returncode == 1 and "no usable CUDA device" in r.stderr and r.stdout == ""


@pytest.mark.gpu
def test_plain_c_samtools_caller_prints_the_references_report(tmp_path):
    """The benchmark's 'RAW SAMTOOLS' reader (flagstats.cpp:490-519, report :577-588) with
    flagstat_loop replaced by FLAGSTAT_cuda_samtools: block by block, whole file, LZ4 and Zstd
    containers."""
    assert os.path.exists(SAMCALL), "oracle/_ref/samtools_caller missing (integration/build_dropin.sh)"
    a = O.synth_hiseqx(0, 5 * 512_000 + 4321, 2, 25_000)
    want = O.samtools_report(O.samtools_loop(a))
    raw = tmp_path / "flags.bin"
    a.tofile(raw)
    lz = tmp_path / "flags.lz4"
    lz.write_bytes(O.write_lz4_container(a))
    cases = [[str(raw)], ["--file", str(raw)], ["--lz4", str(lz)]]
    if O.libzstd() is not None:
        zs = tmp_path / "flags.zst"
        zs.write_bytes(O.write_zstd_container(a, 3))
        cases.append(["--zstd", str(zs)])
    for args in cases:
        r = subprocess.run([SAMCALL] + args, capture_output=True, text=True, timeout=120)
        assert r.returncode == 0, r.stderr
        assert r.stdout == want, (args, r.stdout)
        assert f"{a.size} flags" in r.stderr
