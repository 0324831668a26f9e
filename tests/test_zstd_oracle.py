"""Pins the oracle's Zstd frame decoder (oracle/zstd_oracle.c, restated from RFC 8878)
against the real libzstd of this image (libzstd.so.1 through ctypes, and pyarrow's codec):
frames written by ZSTD_compress at the levels of the reference's table (README.md:148-175:
1..20, plus the writer's default 22, benchmark/flagstats.cpp:192) must decode to the
original bytes, malformed frames must be rejected, and the container walk of
zstd_decompress() (:636-676) must reproduce the column and its counters.  CPU only; the
product's decoder (libflagstats_b200/csrc/zstd_frame.cuh) is tested against this oracle and
against libzstd in tests/test_zstd_frame_host.py (host build) and tests/test_blockfile.py (GPU)."""
import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.skipif(O.libzstd() is None, reason="no libzstd.so.1 runtime")


def _columns():
    rng = np.random.default_rng(5)
    cats = np.array([99, 147, 83, 163, 97, 145, 73, 137, 2113, 77], np.uint16)
    yield "hiseqx", O.synth_hiseqx(0, 512_000, 2, 1000)
    yield "iid", cats[rng.integers(0, 10, 512_000)]
    yield "runs", np.repeat(cats[rng.integers(0, 10, 70_000)], rng.geometric(1 / 8, 70_000))[:512_000]
    yield "uniform12", O.synth_uniform(0, 300_001, 3, 0x0FFF)
    yield "uniform16", O.synth_uniform(0, 200_000, 4, 0xFFFF)
    yield "constant", np.full(512_000, 99, np.uint16)
    yield "period3", np.tile(np.array([99, 147, 83], np.uint16), 100_000)
    yield "tiny", np.array([1, 2, 3], np.uint16)
    yield "one", np.array([7], np.uint16)
    yield "empty", np.zeros(0, np.uint16)


@pytest.mark.parametrize("name,col", list(_columns()))
def test_decoder_matches_libzstd_at_every_level(name, col):
    raw = col.tobytes()
    for level in list(range(1, 23)) + [-1, -5]:
        frame = O.libzstd_compress(raw, level)
        assert O.zstd_decompress(frame, len(raw)) == raw, (name, level)
        assert O.libzstd_decompress(frame, len(raw)) == raw


def test_decoder_matches_pyarrow_codec_and_odd_byte_counts():
    pa = pytest.importorskip("pyarrow")
    rng = np.random.default_rng(9)
    for n in (1, 2, 3, 255, 256, 257, 65_535, 131_071, 131_072, 131_073, 1_024_000 - 1):
        raw = bytes(rng.integers(0, 7, n, dtype=np.uint8))  # low entropy: Huffman literals + sequences
        frame = pa.compress(raw, codec="zstd", asbytes=True)
        assert O.zstd_decompress(frame, n) == raw, n


def test_malformed_frames_are_rejected():
    raw = O.synth_hiseqx(0, 100_000, 1, 0).tobytes()
    frame = O.libzstd_compress(raw, 3)
    with pytest.raises(ValueError):
        O.zstd_decompress(frame[:-5], len(raw))            # truncated
    with pytest.raises(ValueError):
        O.zstd_decompress(frame, len(raw) - 2)             # output too small
    with pytest.raises(ValueError):
        O.zstd_decompress(b"\x00" + frame[1:], len(raw))   # bad magic
    rng = np.random.default_rng(3)
    rejected = 0
    for _ in range(300):  # random corruption: either rejected or (rarely) a different valid stream
        bad = bytearray(frame)
        for _k in range(3):
            bad[int(rng.integers(4, len(bad)))] ^= 1 << int(rng.integers(0, 8))
        try:
            out = O.zstd_decompress(bytes(bad), len(raw))
        except ValueError:
            rejected += 1
            continue
        assert len(out) == len(raw)
    assert rejected > 200


def test_container_walk_reproduces_the_column_and_its_counters():
    col = O.synth_hiseqx(0, 3 * 512_000 + 12_345, 1, 500)
    for level in (1, 9, 20):
        blob = O.write_zstd_container(col, level)
        blocks = list(O.read_zstd_container(blob))
        assert [b.size for b in blocks] == [512_000, 512_000, 512_000, 12_345]
        assert np.array_equal(np.concatenate(blocks), col)
        f = np.zeros(32, np.uint64)
        for b in blocks:
            O.flagstat_simd(b, f)  # one shared counters[32] across blocks, :646,664-665
        assert f.tolist() == O.flagstat_simd(col).tolist()
        # libzstd reading the same container agrees block by block
        ref_blocks = list(O.read_zstd_container(blob, O.libzstd_decompress))
        assert all(np.array_equal(a, b) for a, b in zip(blocks, ref_blocks))


def test_property_random_structured_inputs_round_trip():
    """hypothesis: byte strings built from repeated fragments over small alphabets (the shapes
    that exercise RLE blocks, treeless literals, repeat offsets and all FSE table modes), any
    level: libzstd's frame decodes here to the input."""
    from hypothesis import given, settings, strategies as st

    frag = st.binary(min_size=1, max_size=40)

    @settings(max_examples=120, deadline=None)
    @given(frags=st.lists(frag, min_size=1, max_size=12), picks=st.lists(st.integers(0, 11), min_size=1, max_size=400),
           reps=st.lists(st.integers(1, 60), min_size=1, max_size=400), level=st.sampled_from([-3, 1, 2, 3, 5, 9, 13, 19, 22]))
    def run(frags, picks, reps, level):
        parts = []
        for i, p in enumerate(picks):
            parts.append(frags[p % len(frags)] * reps[i % len(reps)])
        raw = b"".join(parts)
        frame = O.libzstd_compress(raw, level)
        assert O.zstd_decompress(frame, len(raw)) == raw

    run()
