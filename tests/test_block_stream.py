"""Block streaming (BASELINE configs[4]; caller pattern of benchmark/flagstats.cpp:288-358):
1,024,000-byte blocks from a pinned ring, DMA and zero-copy transports, coalesced
submission, short blocks anywhere in the sequence."""
import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("coalesce", [1, 3, 4])
@pytest.mark.parametrize("slots", [2, 4])
def test_stream_modes_match_oracle(cuda_lib, mode, coalesce, slots):
    fs = cuda_lib
    blk = 64_000  # small blocks keep the oracle side fast; the ring logic is size-independent
    a = O.synth_uniform(0, 23 * blk + 777, 5, 0x0FFF)
    # block lengths: mostly full, short ones in the middle and at the end, one empty
    lens = [blk] * 5 + [1234] + [blk] * 7 + [0] + [blk] * 9 + [blk - 1] + [777]
    want = np.zeros(32, np.uint64)
    with fs.BlockStream(0, blk, slots, mode=mode, coalesce=coalesce) as bs:
        off = 0
        for i, ln in enumerate(lens):
            part = a[off:off + ln]
            off += ln
            want += O.flagstat_simd(part)
            if i % 2:
                bs.push(part)
            else:
                slot = bs.acquire()
                slot[:ln] = part
                bs.submit(ln)
        got = bs.finish()
        assert got.tolist() == want.tolist()
        # handle stays usable; accumulator was reset
        bs.push(a[:1000])
        assert bs.finish().tolist() == O.flagstat_simd(a[:1000]).tolist()


def test_stream_reference_block_size_and_selftime(cuda_lib):
    fs = cuda_lib
    n_blocks = 12
    a = O.synth_hiseqx(0, n_blocks * fs.BLOCK_RECORDS, 3, 2000)
    want = O.flagstat_simd(a)
    for mode in (fs.BlockStream.DMA, fs.BlockStream.ZEROCOPY):
        with fs.BlockStream(0, fs.BLOCK_RECORDS, 3, mode=mode, coalesce=4) as bs:
            for i in range(n_blocks):  # 3 groups x 4 blocks: the ring now holds the whole column
                slot = bs.acquire()
                slot[:] = a[i * fs.BLOCK_RECORDS:(i + 1) * fs.BLOCK_RECORDS]
                bs.submit(fs.BLOCK_RECORDS)
            assert bs.finish().tolist() == want.tolist()
            f, sec = bs.selftime(3 * n_blocks)  # three more laps over the same pinned data
            assert f.tolist() == (3 * want).tolist()
            assert sec > 0
