"""The dynamically scheduled twin of the default kernel (flagstat_kernel_dyn.cuh; an A/B variant:
these tests run against an -DFSB_ALL_VARIANTS build and skip on the product library): same counters as
the oracle whatever the number of chunks, the chunk size, the mode, the base alignment, and across
back-to-back / overlapped launches that reuse the per-stream counter slots.  Forced on for short
columns with FLAGSTAT_cuda_set_dynamic(1, cg) -- by default only long columns take it."""
import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _dev(a, off=0):
    import torch
    t = torch.empty(a.size + off + 8, dtype=torch.int16, device="cuda")
    t[off:off + a.size].copy_(torch.from_numpy(a.view(np.int16)))
    return t[off:off + a.size]


@pytest.fixture()
def forced(cuda_lib):
    def force(cg):
        if cuda_lib.lib().FLAGSTAT_cuda_set_dynamic(1, cg) != 0:
            pytest.skip("the dynamically scheduled kernel is an A/B variant (-DFSB_ALL_VARIANTS), "
                        "not compiled into the product library")
    yield force
    cuda_lib.lib().FLAGSTAT_cuda_set_dynamic(-1, 0)


@pytest.mark.parametrize("cg", [1, 2])
def test_golden_cases_through_the_dynamic_kernel(cuda_lib, golden, forced, cg):
    fs = cuda_lib
    forced(cg)
    assert b"flagstat_kernel_dyn" in fs.lib().FLAGSTAT_cuda_kernel_name(0)
    for c in golden["cases"]:
        from tests.helpers import make_input
        a = make_input(c["spec"])
        assert fs.flagstat_u64(_dev(a)).tolist() == c["cuda_expected"], c["name"]
        assert fs.flagstat_u64(_dev(a, 3)).tolist() == c["cuda_expected"], ("base+3", c["name"])


@pytest.mark.parametrize("cg", [1, 2])
def test_lengths_around_chunk_boundaries_all_modes(cuda_lib, forced, cg):
    """0, 1, 2, ... chunks (4096 * cg records each) +- a few records, with left-over vectors and
    ragged records on both sides, far fewer chunks than warps and many more."""
    fs = cuda_lib
    forced(cg)
    chunk = 4096 * cg
    rng = np.random.default_rng(7)
    lens = [0, 1, 7, 8, chunk - 1, chunk, chunk + 1, 2 * chunk + 9, 3 * chunk - 8, 17 * chunk + 4095,
            2368 * chunk - 3, 2368 * chunk + 5, 5000 * chunk + 123, int(rng.integers(10, 20)) * 1_000_003]
    for n in lens:
        for off in (0, 1, 5):
            a = O.synth_uniform(int(rng.integers(0, 1 << 40)), n, int(rng.integers(0, 1 << 30)), 0xFFFF)
            d = _dev(a, off)
            assert fs.flagstat_u64(d).tolist() == O.flagstat_simd(a).tolist(), (n, off)
            assert fs.pospopcnt_u16(d).tolist() == O.pospopcnt(a).tolist(), (n, off)
            want = O.flagstat_simd(a)
            st = O.samtools_loop(a)
            want[0], want[16] = np.uint64(st[2, 0]), np.uint64(st[2, 1])
            assert fs.flagstat_samtools_u64(d).tolist() == want.tolist(), (n, off)


def test_every_record_value_and_dense_mode(cuda_lib, forced):
    fs = cuda_lib
    forced(1)
    ar = np.arange(65536, dtype=np.uint32).astype(np.uint16)
    a = np.concatenate([ar, np.zeros(1, np.uint16), ar, ar[::-1], np.full(3, 0x0FFF, np.uint16), ar[::-1]] * 9)
    assert fs.flagstat_u64(_dev(a)).tolist() == O.flagstat_simd(a).tolist()
    # QC-fail-dense -> clean -> dense again: the detect-free mode is entered and left inside warps
    b = np.concatenate([O.synth_uniform(0, 6_000_000, 7, 0x0FFF), O.synth_hiseqx(0, 6_000_000, 3, 0),
                        O.synth_uniform(1 << 33, 6_000_005, 9, 0xFFFF)])
    assert fs.flagstat_u64(_dev(b, 1)).tolist() == O.flagstat_simd(b).tolist()


def test_back_to_back_and_overlapped_launches_share_the_slot_pair(cuda_lib, forced):
    """Many launches in one stream (plain and with the programmatic-serialization attribute), over
    different columns: every launch must find its counter slot zeroed by the launch that used it
    last, and two streams must not disturb each other."""
    import torch
    fs = cuda_lib
    forced(1)
    cols = [O.synth_uniform(1000 * i, n, 3 + i, 0x0FFF) for i, n in enumerate([5_000_011, 16384 * 8 * 40, 30_000_001, 9_999_999])]
    devs = [_dev(c, i) for i, c in enumerate(cols)]
    wants = [O.flagstat_simd(c) for c in cols]
    torch.cuda.synchronize()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    acc1 = torch.zeros(32, dtype=torch.int64, device="cuda")
    acc2 = torch.zeros(32, dtype=torch.int64, device="cuda")
    reps = 40
    for rep in range(reps):
        for i, d in enumerate(devs):
            fs.flagstat_device(d, out=acc1, stream=s1, overlap=(rep + i) % 2 == 0)
            fs.flagstat_device(devs[(i + 1) % len(devs)], out=acc2, stream=s2, overlap=rep % 3 == 0)
    s1.synchronize()
    s2.synchronize()
    total = sum(wants) * np.uint64(reps)
    assert acc1.cpu().numpy().view(np.uint64).tolist() == total.tolist()
    assert acc2.cpu().numpy().view(np.uint64).tolist() == total.tolist()


def test_product_library_has_the_static_split_only(cuda_lib):
    """The product build must refuse to switch the A/B kernel on (and say so), not silently ignore it."""
    lib = cuda_lib.lib()
    if lib.FLAGSTAT_cuda_set_dynamic(1, 1) == 0:
        lib.FLAGSTAT_cuda_set_dynamic(-1, 0)
        pytest.skip("this is an -DFSB_ALL_VARIANTS build")
    assert lib.FLAGSTAT_cuda_set_dynamic(0, 1) == -2
    assert lib.FLAGSTAT_cuda_set_dynamic(-1, 0) == 0
    assert b"flagstat_kernel_group" in lib.FLAGSTAT_cuda_kernel_name(0)
