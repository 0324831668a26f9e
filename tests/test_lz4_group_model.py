"""The ALGORITHM of csrc/lz4_block_group.cuh (32 sequences per warp step), restated
lane by lane in tests/lz4_group_model.py, against real LZ4 blocks.  CPU only; the
CUDA kernel itself is checked in tests/test_blockfile.py on the GPU."""
import numpy as np
import pytest

from oracle import oracle as O
from tests import lz4_group_model as M

pa = pytest.importorskip("pyarrow")


def _columns():
    rng = np.random.default_rng(11)
    cats = np.array([99, 147, 83, 163, 97, 145, 73, 137, 2113, 77], np.uint16)
    yield "hiseqx", O.synth_hiseqx(0, 60_007, 2, 1000)
    yield "uniform12", O.synth_uniform(0, 20_001, 3, 0x0FFF)
    yield "runs_mean8", np.repeat(cats[rng.integers(0, 10, 6000)], rng.geometric(1 / 8, 6000))
    yield "runs_mean2", np.repeat(cats[rng.integers(0, 10, 20000)], rng.geometric(1 / 2, 20000))
    yield "iid_categories", cats[rng.integers(0, 10, 50_000)]
    yield "long_runs", np.repeat(rng.integers(0, 4096, 60).astype(np.uint16), rng.integers(1, 3000, 60))
    yield "period3", np.tile(np.array([99, 147, 83], np.uint16), 20_000)
    yield "far_matches", np.concatenate([O.synth_uniform(0, 12_000, 9, 0x0FFF)] * 3)  # offsets > 16 KiB
    yield "tiny", np.array([1, 2, 3], np.uint16)


@pytest.mark.parametrize("name,col", list(_columns()))
@pytest.mark.parametrize("ga", [0, 6, 15])
def test_group_algorithm_decodes_real_blocks(name, col, ga):
    raw = col.tobytes()
    for comp in (O.liblz4_compress(raw), O.lz4_compress(raw)):
        stats = {}
        status, out = M.decode(comp, len(raw), ga=ga, seed=ga, stats=stats)
        assert status == len(raw)
        assert out == raw
    if name in ("runs_mean8", "iid_categories", "runs_mean2"):
        # the point of the design: many sequences per group, few dependency rounds
        assert stats["seqs"] / stats["groups"] > 24
        assert stats["rounds"] / stats["groups"] < 6  # pointer-jumping steps, incl. the last (no change)


def test_group_algorithm_rejects_what_the_oracle_rejects():
    raw = np.tile(np.array([99, 147, 83, 163], np.uint16), 5000).tobytes()
    comp = O.liblz4_compress(raw)
    assert M.decode(comp, len(raw))[0] == len(raw)
    assert M.decode(comp[:-3], len(raw))[0] != len(raw)
    assert M.decode(comp, len(raw) - 10)[0] < 0
    bad = bytearray(comp)
    bad[0] = 0x0F
    assert M.decode(bytes(bad), len(raw))[0] < 0
    # simple sequences with an offset that reaches before the block / a zero offset
    assert M.decode(bytes([0x10, 65, 5, 0, 0x10, 66, 1, 0, 0x50, 1, 2, 3, 4, 5]), 64)[0] == -4
    assert M.decode(bytes([0x10, 65, 0, 0, 0x50, 1, 2, 3, 4, 5]), 64)[0] == -4
    # output capacity exceeded inside a group
    assert M.decode(bytes([0x1E, 65, 1, 0, 0x50, 1, 2, 3, 4, 5]), 10)[0] < 0


def test_chained_dependencies_inside_one_group():
    """Hand-made block: every match reads the previous sequence's match output (a chain of
    32 dependent rounds), overlapping copies (offset < length) and offset-1 runs."""
    seqs = bytearray()
    want = bytearray()

    def seq(lit, off, ml):
        assert len(lit) < 15 and 4 <= ml < 274
        seqs.append((len(lit) << 4) | min(ml - 4, 15))
        seqs.extend(lit)
        seqs.extend(bytes([off & 255, off >> 8]))
        if ml >= 19:
            seqs.append(ml - 19)
        want.extend(lit)
        for _ in range(ml):
            want.append(want[-off])

    seq(b"ab", 2, 6)
    for i in range(70):
        seq(bytes([65 + i % 26]), 3 + (i % 4), 4 + (i * 5) % 15)   # reaches into the previous match
    seq(b"", 1, 18)
    seq(b"q", 5, 19)      # extended by a zero byte
    seq(b"", 3, 200)      # overlapping, whole-warp copy, source = the previous extended match
    seq(b"", 150, 273)    # reaches back over several sequences
    seq(b"xyz", 1, 4)
    seq(b"", 7, 18)
    seqs.append(0x50)
    seqs.extend(b"tail!")
    want.extend(b"tail!")
    assert O.lz4_decompress(bytes(seqs), len(want)) == bytes(want)
    for ga in (0, 3):
        stats = {}
        status, out = M.decode(bytes(seqs), len(want), ga=ga, seed=ga, stats=stats)
        assert status == len(want) and out == bytes(want)
    assert stats["rounds"] > stats["groups"]  # chains needed more than one jumping step
