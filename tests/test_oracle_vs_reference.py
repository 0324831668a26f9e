"""Property tests of the oracle restatement against the UNMODIFIED reference
(oracle/_ref, prebuilt by oracle/Makefile from /root/reference).  CPU only.
Skipped only if the prebuilt shim is absent."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from oracle import oracle as O
from tests.helpers import CORE19, CORE20, offset_copy

ref = O.reference()
pytestmark = pytest.mark.skipif(ref is None, reason="oracle/_ref not built (make -C oracle)")


def test_reference_shim_lists_the_expected_kernels():
    names = O.ref_kernels(runnable_only=False)
    assert names[0] == "scalar" and "avx512" in names and "avx512_improved3" in names
    assert O.ref_dispatch_name(10) == "scalar"


@settings(max_examples=60, deadline=None)
@given(n=st.integers(0, 70000), off=st.sampled_from([0, 1, 3, 7]), seed=st.integers(0, 2**32),
       mask=st.sampled_from([0x0FFF, 0x0FFF, 0x03FF, 0x0DFF]))
def test_every_correct_reference_kernel_agrees_with_the_oracle(n, off, seed, mask):
    a = offset_copy(O.synth_uniform(0, n, seed, mask), off)
    want = O.flagstat_simd(a)
    for k in O.REF_CORRECT_KERNELS:
        if not ref.ref_kernel_runnable(k.encode()):
            continue
        got = O.ref_flagstat(k, a).astype(np.uint64)
        assert (got[CORE19] == want[CORE19]).all(), (k, n, off)
        if k in ("sse4", "avx2", "avx512"):
            assert int(got[9]) == int(want[9]), (k, n)
    d = O.ref_flagstats_u16(a).astype(np.uint64)
    assert (d[CORE19] == want[CORE19]).all()


@settings(max_examples=40, deadline=None)
@given(n=st.integers(0, 70000), seed=st.integers(0, 2**32))
def test_bits_12_to_15_are_ignored_like_scalar_and_improved3(n, seed):
    a = O.synth_uniform(0, n, seed, 0xFFFF)
    want = O.flagstat_scalar(a)
    assert O.ref_flagstat("scalar", a).astype(np.uint64).tolist() == want.tolist()
    if ref.ref_kernel_runnable(b"avx512_improved3"):
        assert O.ref_flagstat("avx512_improved3", a).astype(np.uint64).tolist() == want.tolist()


@settings(max_examples=40, deadline=None)
@given(n=st.integers(0, 300000), seed=st.integers(0, 2**32), off=st.sampled_from([0, 1, 5]))
def test_pospopcnt_matches_storm(n, seed, off):
    a = offset_copy(O.synth_uniform(0, n, seed, 0xFFFF), off)
    assert O.ref_pospopcnt(a).astype(np.uint64).tolist() == O.pospopcnt(a).tolist()


def test_accumulate_contract():
    a = O.synth_uniform(0, 5000, 3, 0x0FFF)
    f = O.ref_flagstat("scalar", a)
    f = O.ref_flagstat("scalar", a[:777], f)
    g = O.flagstat_scalar(a)
    g = O.flagstat_scalar(a[:777], g)
    assert f.astype(np.uint64).tolist() == g.tolist()


def test_threaded_wrapper_is_additive():
    a = O.synth_uniform(0, 3_000_001, 11, 0x0FFF)
    one, _ = O.ref_flagstat_mt("scalar", a, 1)
    k = O.best_reference_kernel()
    many, _ = O.ref_flagstat_mt(k, a, 5)
    assert (one[CORE19] == many[CORE19]).all()
    assert (many[CORE20] == O.flagstat_simd(a)[CORE20]).all()
