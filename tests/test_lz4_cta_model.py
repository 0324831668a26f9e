"""The ALGORITHM of csrc/lz4_block_cta.cuh (one CTA per block: window parse with exit maps, tiles with
pointer jumping), restated in tests/lz4_cta_model.py, against real LZ4 blocks from liblz4 and from the
oracle's compressor.  CPU only; the CUDA kernel itself is checked in tests/test_blockfile.py on the GPU."""
import numpy as np
import pytest

from oracle import oracle as O
from tests import lz4_cta_model as M


def _serial_descriptors(block: bytes):
    """the block's sequences the way any serial LZ4 decoder walks them"""
    p, out, d = 0, 0, []
    while True:
        tok = block[p]
        p += 1
        lit = tok >> 4
        if lit == 15:
            while True:
                b = block[p]
                p += 1
                lit += b
                if b != 255:
                    break
        lit_pos = p
        p += lit
        if p >= len(block):
            d.append((out, lit_pos, lit, 0))
            return d, out + lit
        off = block[p] | (block[p + 1] << 8)
        p += 2
        ml = (tok & 15) + 4
        if (tok & 15) == 15:
            while True:
                b = block[p]
                p += 1
                ml += b
                if b != 255:
                    break
        d.append((out, lit_pos, lit, off))
        out += lit + ml


def _columns():
    rng = np.random.default_rng(11)
    cats = np.array([99, 147, 83, 163, 97, 145, 73, 137, 2113, 77], np.uint16)
    yield "hiseqx", O.synth_hiseqx(0, 60_007, 2, 1000)
    yield "uniform12", O.synth_uniform(0, 20_001, 3, 0x0FFF)                 # incompressible: literal runs of KBs
    yield "runs_mean8", np.repeat(cats[rng.integers(0, 10, 6000)], rng.geometric(1 / 8, 6000))
    yield "iid_categories", cats[rng.integers(0, 10, 50_000)]
    yield "long_runs", np.repeat(rng.integers(0, 4096, 60).astype(np.uint16), rng.integers(1, 3000, 60))  # escapes
    yield "period3", np.tile(np.array([99, 147, 83], np.uint16), 20_000)
    yield "far_matches", np.concatenate([O.synth_uniform(0, 12_000, 9, 0x0FFF)] * 3)
    # literal runs of 300 - 1000 bytes between repeats: successors that jump over whole windows (no escape: <= 4 extension bytes)
    parts = []
    for i in range(40):
        parts.append(rng.integers(0, 65536, int(rng.integers(150, 500))).astype(np.uint16))
        parts.append(np.tile(cats[: 3 + i % 5], 40))
    yield "literal_runs_over_windows", np.concatenate(parts)
    yield "tiny", np.array([1, 2, 3], np.uint16)


@pytest.mark.parametrize("name,col", list(_columns()))
def test_window_parse_finds_exactly_the_serial_sequences(name, col):
    raw = col.tobytes()
    for comp in (O.liblz4_compress(raw), O.lz4_compress(raw)):
        want, total = _serial_descriptors(comp)
        stats = {}
        nseq, desc, out = M.parse(comp, len(raw), stats)
        assert nseq == len(want) and out == total == len(raw)
        assert desc == want
        assert stats["sequences"] == len(want) and stats["super_steps"] >= 1


@pytest.mark.parametrize("name,col", list(_columns()))
@pytest.mark.parametrize("ga", [0, 6, 15])
def test_cta_algorithm_decodes_real_blocks(name, col, ga):
    raw = col.tobytes()
    comp = O.liblz4_compress(raw)
    rounds = {}
    for early in (False, True):
        stats = {}
        status, out = M.decode(comp, len(raw), ga=ga, early_root=early, stats=stats)
        assert status == len(raw)
        assert out == raw
        rounds[early] = stats["rounds"]
    assert rounds[True] <= rounds[False]


def test_cta_algorithm_rejects_what_the_oracle_rejects():
    raw = np.tile(np.array([99, 147, 83, 163], np.uint16), 5000).tobytes()
    comp = O.liblz4_compress(raw)
    assert M.decode(comp, len(raw))[0] == len(raw)
    assert M.decode(comp[:-3], len(raw))[0] != len(raw)
    assert M.decode(comp, len(raw) - 10)[0] < 0
    bad = bytearray(comp)
    bad[0] = 0x0F
    assert M.decode(bytes(bad), len(raw))[0] < 0
    # an offset that reaches before the block / a zero offset
    assert M.decode(bytes([0x10, 65, 5, 0, 0x10, 66, 1, 0, 0x50, 1, 2, 3, 4, 5]), 64)[0] == -4
    assert M.decode(bytes([0x10, 65, 0, 0, 0x50, 1, 2, 3, 4, 5]), 64)[0] == -4
    # output capacity exceeded
    assert M.decode(bytes([0x1E, 65, 1, 0, 0x50, 1, 2, 3, 4, 5]), 10)[0] < 0


def test_flag_like_tiles_take_about_five_rounds_and_one_less_with_the_early_exit():
    rng = np.random.default_rng(3)
    cats = np.array([99, 147, 83, 163, 97, 145, 73, 137, 2113, 77], np.uint16)
    raw = cats[rng.integers(0, 10, 60_000)].tobytes()
    comp = O.liblz4_compress(raw)
    a, b = {}, {}
    assert M.decode(comp, len(raw), stats=a)[1] == raw
    assert M.decode(comp, len(raw), early_root=True, stats=b)[1] == raw
    assert 3.0 <= a["rounds"] / a["tiles"] <= 6.0
    assert b["rounds"] < a["rounds"]
