"""The oracle restatement against the golden vectors produced by the unmodified
reference (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from oracle import oracle as O
from tests.helpers import CORE19, CORE20, make_input


def test_golden_has_reference_provenance(golden):
    assert "unmodified" in golden["reference"]
    assert golden["hiseqx"] == {"N": O.HISEQX_N, "M": int(O.oracle().oracle_hiseqx_m())}
    assert len(golden["cases"]) >= 60


def test_hiseqx_multiplier_is_a_bijection():
    from math import gcd
    assert gcd(int(O.oracle().oracle_hiseqx_m()), O.HISEQX_N) == 1


def test_oracle_matches_every_golden_case(golden):
    for c in golden["cases"]:
        a = make_input(c["spec"])
        scalar = O.flagstat_scalar(a)
        simd = O.flagstat_simd(a)
        # FLAGSTAT_scalar writes nothing outside CORE19: compare all 32 slots
        assert scalar.tolist() == c["scalar"], c["name"]
        assert simd.tolist() == c["cuda_expected"], c["name"]
        assert O.flagstat_maskselect(a).tolist() == c["cuda_expected"], c["name"]
        assert O.numpy_flagstat(a).tolist() == c["cuda_expected"], c["name"]
        assert O.pospopcnt(a).tolist() == c["pospopcnt"], c["name"]
        if c["valid_sam"]:
            d = np.array(c["dispatch"], np.uint64)
            assert (simd[CORE19] == d[CORE19]).all(), c["name"]
            if c["dispatch_kernel"] != "scalar":  # SIMD kernels add n_pass to slot 9
                assert int(d[9]) == int(simd[9]), c["name"]


def test_kat_a_closed_form(golden):
    c = next(c for c in golden["cases"] if c["name"].startswith("KAT-A"))
    # SURVEY.md section 8c, KAT A
    assert c["scalar"][:16] == [0, 0, 1024, 0, 0, 0, 128, 128, 1024, 0, 1024, 512, 64, 64, 64, 0]
    assert c["scalar"][16:] == [0, 0, 1024, 0, 0, 0, 128, 128, 1024, 2048, 1024, 512, 64, 64, 64, 0]
    assert c["cuda_expected"][9] == 2048
    assert c["pospopcnt"] == [2048] * 12 + [0] * 4


def test_kat_c_accumulate(golden):
    a = make_input({"gen": "arange", "n": 4096})
    f = O.flagstat_simd(a)
    f = O.flagstat_simd(a[:2048], f)
    assert f.tolist() == golden["kat_c"]["cuda_expected"]
    assert int(f[25]) == 3072 and int(f[9]) == 3072 and int(f[2]) == 1536


def test_u32_entry_wraps_like_the_reference():
    a = make_input({"gen": "uniform", "n": 1000, "seed": 1, "mask": 0x0FFF})
    f = np.full(32, 0xFFFFFFFF, np.uint32)
    O.flagstat_simd_u32(a, f)
    want = (O.flagstat_simd(a) + np.uint64(0xFFFFFFFF)) & np.uint64(0xFFFFFFFF)
    assert f.astype(np.uint64).tolist() == want.tolist()


def test_mask_select_truth_table_matches_paper_scripts():
    """paper/scripts/mask_data.py:29-46 and expand_data.py:3-10, restated
    independently here, against oracle_mask_select for all 4096 valid flags."""
    for v in range(4096):
        p, pp, u, mu = v & 1, (v >> 1) & 1, (v >> 2) & 1, (v >> 3) & 1
        x = v | ((p & pp & (1 - u)) << 12) | ((p & mu & (1 - u)) << 13) | \
            ((p & (1 - u) & (1 - mu)) << 14)
        keep = 0x4 | 0x200 | 0x400
        if v & 0x100:
            keep |= 0x100
        elif v & 0x800:
            keep |= 0x800
        elif p:
            keep |= 0x7000 | 0x40 | 0x80
        assert O.mask_select(v) == (x & keep), v
        assert O.mask_select(v | 0xF000) == (x & keep), v  # bits 12-15 ignored


@pytest.mark.parametrize("start,n", [(0, 4096), (123456789, 5000), (O.HISEQX_N - 100, 300)])
def test_generators_are_pure_functions_of_the_global_index(start, n):
    whole = O.synth_uniform(start, n, 9, 0x0FFF)
    parts = np.concatenate([O.synth_uniform(start, 1000, 9, 0x0FFF),
                            O.synth_uniform(start + 1000, n - 1000, 9, 0x0FFF)]) if n > 1000 else whole
    assert (whole == parts).all()
    h = O.synth_hiseqx(start, n, 0, 0)
    h2 = np.concatenate([O.synth_hiseqx(start, 100, 0, 0), O.synth_hiseqx(start + 100, n - 100, 0, 0)])
    assert (h == h2).all()
    assert ((O.synth_hiseqx(start, n, 5, 50000) & ~np.uint16(0x200)) == h).all()


def test_kat_e_full_hiseqx_column(golden):
    """824,541,892 records; every FLAG-derivable line of README.md:179-191."""
    e = golden["kat_e"]
    a = make_input(e["spec"])
    vals, cnts = np.unique(a, return_counts=True)
    assert vals.tolist() == e["category_values"] and cnts.tolist() == e["category_counts"]
    f = O.numpy_flagstat(a)
    assert f.tolist() == e["cuda_expected"]
    r = e["readme"]
    assert int(f[9]) + int(f[25]) == r["total"] == a.size
    assert r["total"] - int(f[2]) == r["mapped"]
    assert int(f[6]) + int(f[7]) == r["paired"]
    assert (int(f[12]), int(f[14]), int(f[13])) == (r["properly_paired"], r["both_mapped"], r["singletons"])
    # C restatement on each of the 8 shards sums to the whole (additivity)
    tot = np.zeros(32, np.uint64)
    for s in e["shards8"][:2]:
        g = O.flagstat_simd(a[s["start"]: s["start"] + s["n"]])
        assert g.tolist() == s["cuda_expected"]
    for s in e["shards8"]:
        tot += np.array(s["cuda_expected"], np.uint64)
    assert tot[CORE20].tolist() == f[CORE20].tolist()


def test_kat_16g_is_twenty_periods_plus_a_prefix(golden):
    """BASELINE configs[3] (2^34 records): the generator is periodic, so the golden answer is
    20 x KAT-E + the counters of a 689,031,344-record prefix; the restatement recomputes the
    prefix in 2^26-record slices (the GPU test counts all 2^34 records on one device)."""
    k = golden["kat_16g"]
    n, per, rem = k["spec"]["n"], k["periods"], k["prefix"]
    assert n == 1 << 34 and per * O.HISEQX_N + rem == n and 0 < rem < O.HISEQX_N
    f = np.zeros(32, np.uint64)
    pair_all = 0
    step = 1 << 26
    for lo in range(0, rem, step):
        a = O.synth_hiseqx(lo, min(step, rem - lo))
        O.flagstat_simd(a, f)
        pair_all += int(O.samtools_loop(a)[2, 0])
    assert f.tolist() == k["prefix_cuda_expected"]
    e = np.array(golden["kat_e"]["cuda_expected"], np.uint64)
    assert (np.uint64(per) * e + f).tolist() == k["cuda_expected"]
    assert per * golden["kat_e"]["samtools"][2][0] + pair_all == k["n_pair_all"]
