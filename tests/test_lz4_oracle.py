"""Pins the oracle's LZ4 block restatement (oracle/lz4_oracle.c) against a real
liblz4 (pyarrow's lz4_raw codec) in both directions, and the container walk
against the plain column.  CPU only."""
import numpy as np
import pytest

from oracle import oracle as O

pa = pytest.importorskip("pyarrow")


def _columns():
    rng = np.random.default_rng(5)
    yield "hiseqx", O.synth_hiseqx(0, 300_007, 2, 1000)
    yield "uniform12", O.synth_uniform(0, 70_001, 3, 0x0FFF)
    yield "runs", np.repeat(rng.integers(0, 4096, 500).astype(np.uint16), rng.integers(1, 3000, 500))
    yield "constant", np.full(600_000, 99, np.uint16)
    yield "period3", np.tile(np.array([99, 147, 83], np.uint16), 100_000)
    yield "tiny", np.array([1, 2, 3], np.uint16)
    yield "empty", np.zeros(0, np.uint16)


@pytest.mark.parametrize("name,col", list(_columns()))
def test_decoder_matches_liblz4(name, col):
    raw = col.tobytes()
    comp = O.liblz4_compress(raw)
    assert O.lz4_decompress(comp, len(raw)) == raw


@pytest.mark.parametrize("name,col", list(_columns()))
def test_encoder_output_decodes_with_liblz4_and_oracle(name, col):
    raw = col.tobytes()
    comp = O.lz4_compress(raw)
    assert O.lz4_decompress(comp, len(raw)) == raw
    if raw:
        assert O.liblz4_decompress(comp, len(raw)) == raw


def test_malformed_blocks_are_rejected():
    raw = np.tile(np.array([99, 147, 83, 163], np.uint16), 5000).tobytes()
    comp = bytearray(O.liblz4_compress(raw))
    with pytest.raises(ValueError):
        O.lz4_decompress(bytes(comp[:-3]), len(raw))       # truncated
    with pytest.raises(ValueError):
        O.lz4_decompress(bytes(comp), len(raw) - 10)       # output too small
    bad = bytearray(comp)
    bad[0] = 0x0F  # no literals, then a match with nothing to copy from
    with pytest.raises(ValueError):
        O.lz4_decompress(bytes(bad), len(raw))


def test_container_walk_reproduces_the_column():
    col = O.synth_hiseqx(0, 3 * 512_000 + 12_345, 1, 500)
    for compressor in (O.liblz4_compress, O.lz4_compress):
        blob = O.write_lz4_container(col, compressor=compressor)
        blocks = list(O.read_lz4_container(blob))
        assert [b.size for b in blocks] == [512_000, 512_000, 512_000, 12_345]
        assert np.array_equal(np.concatenate(blocks), col)
        f = np.zeros(32, np.uint64)
        for b in blocks:
            O.flagstat_simd(b, f)
        assert f.tolist() == O.flagstat_simd(col).tolist()
