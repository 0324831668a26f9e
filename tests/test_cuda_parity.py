"""Parity of the CUDA path against the oracle and the golden vectors of the
unmodified reference.  Everything here goes through the C ABI
(include/flagstats_cuda.h) via libflagstats_b200/_capi.py and needs a B200.

Bar: bit-exact on CORE19 + slot 9, every other slot untouched."""
import ctypes as C

import numpy as np
import pytest

from oracle import oracle as O
from tests.helpers import CORE20, UNTOUCHED, make_input, offset_copy

pytestmark = pytest.mark.gpu


def _torch():
    import torch
    assert torch.cuda.is_available()
    return torch


def _dev(a, off=0):
    """numpy uint16 -> CUDA int16 tensor whose base is `off` records past a
    256-byte-aligned allocation."""
    torch = _torch()
    t = torch.empty(a.size + off + 8, dtype=torch.int16, device="cuda")
    t[off:off + a.size].copy_(torch.from_numpy(a.view(np.int16)))
    v = t[off:off + a.size]
    assert a.size == 0 or (v.data_ptr() - 2 * off) % 256 == 0
    return v


def _select_variant(fs, variant):
    """Returns the previous variant; skips the test when this build does not contain `variant`
    (the product library ships 0, 1, 3 and 8; the rest need -DFSB_ALL_VARIANTS)."""
    prev = fs.lib().FLAGSTAT_cuda_set_variant(variant)
    if prev < 0:
        assert variant not in (0, 1, 3, 8), "a product variant is missing from the build"
        pytest.skip(f"kernel variant {variant} is an A/B variant, not compiled into the product library")
    return prev


@pytest.mark.parametrize("variant", [0, 1, 2, 3, 4, 5, 6, 7, 8])
def test_golden_cases_host_and_device_pointers(cuda_lib, golden, variant):
    fs = cuda_lib
    prev = _select_variant(fs, variant)
    try:
        for c in golden["cases"]:
            a = make_input(c["spec"])
            want = c["cuda_expected"]
            got_h = fs.flagstat_u64(a)
            assert got_h.tolist() == want, ("host", c["name"])
            got_d = fs.flagstat_u64(_dev(a))
            assert got_d.tolist() == want, ("device", c["name"])
    finally:
        fs.lib().FLAGSTAT_cuda_set_variant(prev)


@pytest.mark.parametrize("variant", [0, 1, 2, 7, 8])
def test_every_record_value_in_both_register_halves(cuda_lib, variant):
    """All 65536 FLAG words, once at even and once at odd record positions (low and
    high half of the packed 32-bit register), and next to every neighbour class."""
    fs = cuda_lib
    ar = np.arange(65536, dtype=np.uint32).astype(np.uint16)
    a = np.concatenate([ar, np.zeros(1, np.uint16), ar, ar[::-1], np.full(3, 0x0FFF, np.uint16), ar[::-1]])
    want = O.flagstat_simd(a).tolist()
    prev = _select_variant(fs, variant)
    try:
        assert fs.flagstat_u64(_dev(a)).tolist() == want
        assert fs.flagstat_u64(_dev(a, 1)).tolist() == want
    finally:
        fs.lib().FLAGSTAT_cuda_set_variant(prev)


@pytest.mark.parametrize("off", [1, 3, 7, 5])
def test_unaligned_bases(cuda_lib, golden, off):
    fs = cuda_lib
    for c in golden["cases"]:
        if c["spec"]["n"] > 200000:
            continue
        a = make_input(c["spec"])
        assert fs.flagstat_u64(_dev(a, off)).tolist() == c["cuda_expected"], (off, c["name"])
        assert fs.flagstat_u64(offset_copy(a, off)).tolist() == c["cuda_expected"], (off, c["name"])
        assert fs.pospopcnt_u16(_dev(a, off)).tolist() == c["pospopcnt"], (off, c["name"])


def test_reference_signature_accumulates_wraps_and_leaves_other_slots(cuda_lib, golden):
    fs = cuda_lib
    c = next(c for c in golden["cases"] if c["name"].startswith("KAT-D"))
    a = make_input(c["spec"])
    # KAT-C style accumulation through the uint32 entry (libflagstats.h:2970)
    f = np.zeros(32, np.uint32)
    fs.flagstat_u32(a, f)
    assert f.astype(np.uint64).tolist() == c["cuda_expected"]
    sentinel = np.full(32, 0xFFFFFFF0, np.uint32)
    fs.flagstat_u32(a, sentinel)
    want = (np.array(c["cuda_expected"], np.uint64) + np.uint64(0xFFFFFFF0)) & np.uint64(0xFFFFFFFF)
    assert sentinel.astype(np.uint64).tolist() == want.tolist()
    assert all(int(sentinel[i]) == 0xFFFFFFF0 for i in UNTOUCHED)
    # KAT-C exactly
    ka = make_input({"gen": "arange", "n": 4096})
    g = np.zeros(32, np.uint64)
    fs.flagstat_u64(ka, g)
    fs.flagstat_u64(ka[:2048], g)
    assert g.tolist() == golden["kat_c"]["cuda_expected"]


def test_pyflagstats_compatible_dict(cuda_lib, golden):
    fs = cuda_lib
    c = next(c for c in golden["cases"] if c["name"].startswith("KAT-D"))
    a = make_input(c["spec"])
    d = fs.flagstats(a)
    want = fs.counters_to_dict(np.array(c["cuda_expected"], np.uint64).astype(np.uint32), a.size)
    assert d["n_values"] == a.size
    assert {k: int(v) for k, v in d["passed"].items()} == {k: int(v) for k, v in want["passed"].items()}
    assert {k: int(v) for k, v in d["failed"].items()} == {k: int(v) for k, v in want["failed"].items()}
    with pytest.raises(ValueError):
        fs.flagstats(list(a[:10]))
    with pytest.raises(ValueError):
        fs.flagstats(a.astype(np.int32))
    assert fs.flagstats(a[::2])["n_values"] == a[::2].size  # non-contiguous gets fixed


def test_pospopcnt_golden(cuda_lib, golden):
    fs = cuda_lib
    for c in golden["cases"]:
        a = make_input(c["spec"])
        assert fs.pospopcnt_u16(a).tolist() == c["pospopcnt"], c["name"]
        out = np.full(16, 0xDEADBEEF, np.uint32)  # must be overwritten (libalgebra.h:3498)
        fs.check(fs.lib().POSPOPCNT_cuda_u16(a.ctypes.data, a.size,
                                             out.ctypes.data_as(C.POINTER(C.c_uint32))), "pospopcnt")
        assert out.astype(np.uint64).tolist() == [x & 0xFFFFFFFF for x in c["pospopcnt"]]


def test_device_generators_equal_oracle_generators(cuda_lib):
    from libflagstats_b200 import synth
    for start, n in [(0, 100003), (1_000_000_007, 70001), (O.HISEQX_N - 1000, 5000)]:
        u = synth.uniform_device(n, start, 5, 0x0FFF).cpu().numpy().view(np.uint16)
        assert (u == O.synth_uniform(start, n, 5, 0x0FFF)).all()
        h = synth.hiseqx_device(n, start, 3, 20000).cpu().numpy().view(np.uint16)
        assert (h == O.synth_hiseqx(start, n, 3, 20000)).all()


def test_hypothesis_style_random_lengths_offsets(cuda_lib):
    fs = cuda_lib
    rng = np.random.default_rng(1234)
    for _ in range(60):
        n = int(rng.choice([rng.integers(0, 300), rng.integers(0, 20000), rng.integers(0, 3_000_000)]))
        off = int(rng.integers(0, 8))
        mask = int(rng.choice([0x0FFF, 0xFFFF, 0x03FF]))
        a = O.synth_uniform(int(rng.integers(0, 1 << 40)), n, int(rng.integers(0, 1 << 30)), mask)
        want = O.flagstat_simd(a)
        assert fs.flagstat_u64(_dev(a, off)).tolist() == want.tolist(), (n, off, mask)
        assert fs.pospopcnt_u16(_dev(a, off)).tolist() == O.pospopcnt(a).tolist(), (n, off, mask)


def test_inmemory_config_100m_uniform(cuda_lib):
    """BASELINE configs[0]: 100 M U(0,4095) words, bit-exact vs the oracle, plus
    permutation invariance and shard additivity."""
    fs = cuda_lib
    from libflagstats_b200 import synth
    torch = _torch()
    n = 100_000_000
    d = synth.uniform_device(n, 0, 0, 0x0FFF)
    got = fs.flagstat_u64(d)
    host = d.cpu().numpy().view(np.uint16)
    assert (host[:100000] == O.synth_uniform(0, 100000, 0, 0x0FFF)).all()
    assert got.tolist() == O.numpy_flagstat(host).tolist()
    k = 41_234_567
    parts = fs.flagstat_u64(d[:k])
    fs.flagstat_u64(d[k:], parts)
    assert parts.tolist() == got.tolist()
    assert fs.flagstat_u64(torch.flip(d, dims=[0]).contiguous()).tolist() == got.tolist()
    assert fs.flagstat_u64(host).tolist() == got.tolist()  # pageable host pointer, chunked staging


@pytest.mark.parametrize("variant", [0, 8])
def test_dense_mode_transitions(cuda_lib, variant):
    """The default kernel (variant 0) switches its per-batch OR-detect off after a group of
    batches that all fed the QC-fail counter and probes again 7 groups later (variant 8, the
    round-1 default, has no such mode and runs as the cross-check).  A column whose phases are long enough for every CTA to enter the mode, run it
    over QC-clean data, leave it and enter it again must still count exactly; with 50 000
    records of bits 12..15 garbage sprinkled in (the fp16 gating needs clean high bits)."""
    fs = cuda_lib
    from libflagstats_b200 import synth
    torch = _torch()
    phase = 120_000_000
    d = torch.empty(3 * phase + 5, dtype=torch.int16, device="cuda")
    synth.uniform_device(phase, 0, 7, 0x0FFF, out=d[:phase])
    synth.hiseqx_device(phase, 0, 3, 0, out=d[phase:2 * phase])
    synth.uniform_device(phase + 5, 1 << 33, 9, 0xFFFF, out=d[2 * phase:])
    torch.cuda.synchronize()
    host = d.cpu().numpy().view(np.uint16)
    want = O.flagstat_simd(host)
    head = O.flagstat_simd(host[:3])
    prev = _select_variant(fs, variant)
    try:
        assert fs.flagstat_u64(d).tolist() == want.tolist()
        assert fs.flagstat_u64(d[3:]).tolist() == (want - head).tolist()  # unaligned base
    finally:
        fs.lib().FLAGSTAT_cuda_set_variant(prev)


def test_kat_e_full_hiseqx_on_device(cuda_lib, golden):
    """BASELINE configs[1]: 824,541,892 HiSeqX-shaped FLAGs."""
    fs = cuda_lib
    from libflagstats_b200 import synth
    e = golden["kat_e"]
    d = synth.hiseqx_device(O.HISEQX_N)
    got = fs.flagstat_u64(d)
    assert got.tolist() == e["cuda_expected"]
    assert fs.pospopcnt_u16(d).tolist() == e["pospopcnt"]
    rep = fs.samtools_report(got)
    assert "824541892 + 0 in total" in rep and "805383403 + 0 mapped (97.68% : N/A)" in rep
    assert "781085884 + 0 properly paired (95.35% : N/A)" in rep
    assert "2038885 + 0 singletons (0.25% : N/A)" in rep
    for s in e["shards8"]:
        assert fs.flagstat_u64(d[s["start"]: s["start"] + s["n"]]).tolist() == s["cuda_expected"]
    # QC-fail knob exercises the second counter on the same column: pass + fail == clean
    q = synth.hiseqx_device(50_000_000, 0, 9, 30000)
    gq = fs.flagstat_u64(q)
    clean = fs.flagstat_u64(d[:50_000_000])
    assert (gq[:16] + gq[16:]).tolist() == (clean[:16] + clean[16:]).tolist()
    assert int(gq[25]) > 0


def test_more_than_2_pow_32_records_and_epoch_flush(cuda_lib, golden):
    """6 x 824,541,892 = 4,947,251,352 records (> 2^32, 9.9 GB): the generator is
    periodic, so the answer is exactly 6 x KAT-E; per-thread input exceeds one
    counter epoch, exercising the in-kernel flush; u64 length entry."""
    fs = cuda_lib
    from libflagstats_b200 import synth
    n = 6 * O.HISEQX_N
    assert n > 2 ** 32
    d = synth.hiseqx_device(n)
    got = fs.flagstat_u64(d)
    assert got.tolist() == [6 * x for x in golden["kat_e"]["cuda_expected"]]
    prev = fs.lib().FLAGSTAT_cuda_set_ctas_per_sm(1)
    try:
        assert fs.flagstat_u64(d).tolist() == got.tolist()
    finally:
        fs.lib().FLAGSTAT_cuda_set_ctas_per_sm(prev)


def test_baseline_config3_16g_records_on_one_gpu(cuda_lib, golden):
    """BASELINE configs[3] at its full size on ONE B200: 2^34 HiSeqX-shaped records (34.4 GB of
    HBM) = 20 periods of the generator + a 689,031,344-record prefix; the expected counters were
    made with the reference (tests/golden/make_golden.py, kat_16g).  Also the 8 range shards of
    the multi-GPU layout, counted one after the other, must add up to the same answer."""
    fs = cuda_lib
    torch = _torch()
    from libflagstats_b200 import sharded, synth
    k = golden["kat_16g"]
    n = k["spec"]["n"]
    free, _total = torch.cuda.mem_get_info()
    if free < 2 * n + (2 << 30):
        pytest.skip("needs 36 GB of free device memory")
    d = synth.hiseqx_device(n)
    got = fs.flagstat_u64(d)
    assert got.tolist() == k["cuda_expected"]
    sam = fs.flagstat_samtools_u64(d)
    assert int(sam[0]) == k["n_pair_all"] and int(sam[16]) == 0
    assert sam[CORE20].tolist() == got[CORE20].tolist()
    acc = np.zeros(32, np.uint64)
    for r in range(8):
        lo, hi = sharded.shard_range(n, 8, r)
        fs.flagstat_u64(d[lo:hi], acc)
    assert acc.tolist() == got.tolist()
    del d
    torch.cuda.empty_cache()


def test_async_device_entry_and_stream_blocks(cuda_lib):
    fs = cuda_lib
    torch = _torch()
    a = O.synth_uniform(0, 5_300_123, 21, 0x0FFF)
    want = O.numpy_flagstat(a)
    d = _dev(a)
    out = fs.flagstat_device(d)
    fs.flagstat_device(d, out=out)  # accumulates
    torch.cuda.synchronize()
    assert (out.cpu().numpy().view(np.uint64) == 2 * want).all()
    pp = fs.flagstat_device(d, pospopcnt=True)
    torch.cuda.synchronize()
    assert pp.cpu().numpy().view(np.uint64).tolist() == O.pospopcnt(a).tolist()
    # 1,024,000-byte blocks, the caller pattern of benchmark/flagstats.cpp:288-358
    with fs.BlockStream(0, fs.BLOCK_RECORDS, 4) as bs:
        for lo in range(0, a.size, fs.BLOCK_RECORDS):
            blk = a[lo:lo + fs.BLOCK_RECORDS]
            if (lo // fs.BLOCK_RECORDS) % 2:
                bs.push(blk)
            else:
                slot = bs.acquire()
                slot[:blk.size] = blk
                bs.submit(blk.size)
        got = bs.finish()
        assert got.tolist() == want.tolist()
        bs.push(a[:1000])  # handle stays usable, accumulator was reset
        assert bs.finish().tolist() == O.flagstat_simd(a[:1000]).tolist()


def test_multi_device_entry_and_errors(cuda_lib):
    fs = cuda_lib
    a = O.synth_uniform(0, 3_000_001, 2, 0x0FFF)
    want = O.numpy_flagstat(a)
    f = np.zeros(32, np.uint64)
    fs.check(fs.lib().FLAGSTAT_cuda_multi_u64(a.ctypes.data, a.size,
                                              f.ctypes.data_as(C.POINTER(C.c_uint64)), 0), "multi")
    assert f.tolist() == want.tolist()
    # odd address -> EINVAL, flags untouched
    g = np.full(32, 7, np.uint64)
    rc = fs.lib().FLAGSTAT_cuda_u64(a.ctypes.data + 1, 10, g.ctypes.data_as(C.POINTER(C.c_uint64)))
    assert rc == -2 and (g == 7).all()
    assert fs.lib().FLAGSTAT_cuda_launch_count() > 0


@pytest.mark.parametrize("n,off", [(8_388_608, 0), (20_000_003, 1), (37_123_457, 3)])
def test_pageable_host_arrays_take_the_threaded_staging_path(cuda_lib, n, off):
    """numpy memory is pageable: arrays >= 16 MiB go through run_pageable() (T threads copy
    slices into pinned slots; one DMA + launch per slice).  Same counters as the oracle for
    flagstat, the reference's uint32 signature and pospopcnt; ragged lengths, odd bases."""
    fs = cuda_lib
    a = offset_copy(O.synth_hiseqx(off, n, 5, 20000), off)
    launches = fs.lib().FLAGSTAT_cuda_launch_count()
    want = O.flagstat_simd(a)
    assert fs.flagstat_u64(a).tolist() == want.tolist()
    assert fs.lib().FLAGSTAT_cuda_launch_count() - launches >= 2  # sliced, not one chunk
    f = fs.flagstat_u32(a)
    assert f.astype(np.uint64)[CORE20].tolist() == (want[CORE20] & 0xFFFFFFFF).tolist()
    assert fs.pospopcnt_u16(a).tolist() == O.pospopcnt(a).tolist()
