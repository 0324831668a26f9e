"""FLAG ingest on the GPU (benchmark/utility.cpp:29-32): text lines ->
uint16 column with getline + atoi semantics, checked against the C library's atoi."""
import numpy as np
import pytest

from oracle import oracle as O

ODD_LINES = [b"99", b"147", b"", b"  42", b"\t7", b"+12", b"-1", b"12abc", b"abc", b"99\r", b"70000",
             b"00000000000099", b"4095", b"0", b" ", b"- 5", b"65535", b"1 2", b"2113"]


def test_oracle_ingest_semantics():
    got = O.ingest_text(b"\n".join(ODD_LINES) + b"\n")
    assert got.tolist() == [99, 147, 0, 42, 7, 12, 65535, 12, 0, 99, 70000 & 0xFFFF, 99, 4095, 0, 0, 0,
                            65535, 1, 2113]


@pytest.mark.gpu
@pytest.mark.parametrize("trailing_newline", [True, False])
def test_gpu_ingest_matches_atoi(cuda_lib, trailing_newline):
    from libflagstats_b200 import blockfile

    rng = np.random.default_rng(11)
    col = O.synth_hiseqx(0, 300_001, 5, 20000)
    lines = [str(int(v)).encode() for v in col]
    # sprinkle the odd lines all over, including at tile boundaries
    for k in range(0, len(lines), 997):
        lines[k] = ODD_LINES[(k // 997) % len(ODD_LINES)]
    text = b"\n".join(lines) + (b"\n" if trailing_newline else b"")
    want = O.ingest_text(text)
    assert want.size == len(lines)
    got, flags = blockfile.ingest_text(text, with_flags=True)
    assert got.tolist() == want.tolist()
    assert flags.tolist() == O.flagstat_simd(want).tolist()
    # misaligned host buffer start
    arr = np.frombuffer(b"x" + text, dtype=np.uint8)[1:]
    assert blockfile.ingest_text(arr).tolist() == want.tolist()
    assert rng is not None


@pytest.mark.gpu
def test_gpu_ingest_edge_cases(cuda_lib):
    from libflagstats_b200 import blockfile

    for text in (b"", b"\n", b"\n\n\n", b"7", b"7\n", b"\n7", b"99\n147", b" \n \n", b"1\n" * 5000 + b"2"):
        assert blockfile.ingest_text(text).tolist() == O.ingest_text(text).tolist(), text[:20]
