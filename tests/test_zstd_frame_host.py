"""The PRODUCT's Zstd frame decoder (libflagstats_b200/csrc/zstd_frame.cuh, __host__ __device__)
compiled for the host with g++ -fsanitize=address,undefined and held to the real libzstd
(libzstd.so.1) and to the oracle: every compression level of the reference's table on
FLAG-shaped and adversarial columns, exact-size buffers, corrupted frames.  The GPU tests
(tests/test_blockfile.py, -m gpu) run the same source on the device.  CPU only."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(O.libzstd() is None or shutil.which("g++") is None,
                                reason="needs libzstd.so.1 and g++")


@pytest.fixture(scope="module")
def host(tmp_path_factory):
    so = tmp_path_factory.mktemp("zstd_host") / "libzstd_frame_host.so"
    cmd = ["g++", "-std=c++17", "-O1", "-g", "-fPIC", "-shared", "-Wall", "-Wextra", "-Werror",
           "-fsanitize=undefined", "-fno-sanitize-recover=all",
           "-I", os.path.join(ROOT, "libflagstats_b200", "csrc"), "-x", "c++",
           os.path.join(ROOT, "tests", "native", "zstd_frame_host.cpp"), "-o", str(so)]
    subprocess.check_call(cmd)
    lib = C.CDLL(str(so))
    for name in ("zstd_frame_host", "zstd_frame_host_v2"):
        getattr(lib, name).argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64]
        getattr(lib, name).restype = C.c_int64
    lib.zstd_frame_work_bytes.restype = C.c_uint64
    lib.zstd_frame_tables_bytes.restype = C.c_uint64
    return lib


def _decode(lib, frame: bytes, cap: int):
    """Both versions of the product decoder: decode_frame (one thread copies as it goes) and parse_frame +
    apply_descriptors (entropy stage into descriptors, what the library runs on the device with the copies
    done by l4_copy).  They must agree with each other on every input; the first one's result is returned."""
    src = (C.c_ubyte * max(len(frame), 1)).from_buffer_copy(frame or b"\0")
    dst = (C.c_ubyte * max(cap, 1))()
    r = lib.zstd_frame_host(src, len(frame), dst, cap)
    out = bytes(dst[: max(r, 0)])
    dst2 = (C.c_ubyte * max(cap, 1))()
    r2 = lib.zstd_frame_host_v2(src, len(frame), dst2, cap)
    assert (r2 < 0) == (r < 0), (r, r2)
    if r >= 0:
        assert r2 == r and bytes(dst2[:r]) == out
    return r, out


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
    yield "empty", np.zeros(0, np.uint16)


def test_workspace_is_what_the_library_allocates_per_frame(host):
    assert 130_000 < host.zstd_frame_work_bytes() < 160_000
    # second version: the tables of one frame live in shared memory, four frames per CTA, three CTAs per SM
    assert 3 * (host.zstd_frame_tables_bytes() * 4 + 1024) <= 228 * 1024


@pytest.mark.parametrize("name,col", list(_columns()))
def test_product_decoder_matches_libzstd_at_every_level(host, name, col):
    raw = col.tobytes()
    for level in list(range(1, 23)) + [-1, -5]:
        frame = O.libzstd_compress(raw, level)
        r, out = _decode(host, frame, len(raw))
        assert r == len(raw) and out == raw, (name, level, r)


def test_product_decoder_on_low_entropy_bytes_and_block_boundaries(host):
    rng = np.random.default_rng(9)
    for n in (1, 2, 3, 255, 256, 257, 65_535, 131_071, 131_072, 131_073, 1_024_000 - 1, 1_024_000):
        raw = bytes(rng.integers(0, 7, n, dtype=np.uint8))
        for level in (1, 3, 19):
            frame = O.libzstd_compress(raw, level)
            r, out = _decode(host, frame, n)
            assert r == n and out == raw, (n, level, r)


def test_product_decoder_agrees_with_the_oracle_on_corrupted_frames(host):
    """Same verdict (accepted / rejected) and same bytes as oracle/zstd_oracle.c for every
    corrupted frame: the two decoders were written separately from the same RFC."""
    raw = O.synth_hiseqx(0, 100_000, 1, 0).tobytes()
    frame = O.libzstd_compress(raw, 3)
    rng = np.random.default_rng(3)
    rejected = 0
    for t in range(400):
        bad = bytearray(frame)
        for _k in range(1 + t % 3):
            bad[int(rng.integers(4, len(bad)))] ^= 1 << int(rng.integers(0, 8))
        cut = len(bad) if t % 4 else int(rng.integers(0, len(bad) + 1))
        bad = bytes(bad[:cut])
        r, out = _decode(host, bad, len(raw))
        try:
            want = O.zstd_decompress(bad, len(raw))
        except ValueError:
            want = None
        if want is None:
            assert r != len(raw), t
            rejected += 1
        else:
            assert r == len(raw) and out == want, t
    assert rejected > 250
    assert _decode(host, frame, len(raw) - 2)[0] < 0  # output too small


def test_product_decoder_under_asan_and_ubsan_fuzz(tmp_path):
    exe = tmp_path / "zstd_frame_fuzz"
    build = subprocess.run(
        ["g++", "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=all",
         "-fno-omit-frame-pointer", "-Wall", "-Wextra", "-Werror",
         "-I", os.path.join(ROOT, "libflagstats_b200", "csrc"), "-x", "c++",
         os.path.join(ROOT, "tests", "native", "zstd_frame_fuzz.cpp"), "-o", str(exe), "-l:libzstd.so.1"],
        capture_output=True, text=True, timeout=300)
    if build.returncode != 0 and "sanitize" in build.stderr and "cannot find" in build.stderr:
        pytest.skip("toolchain has no sanitizer runtime")
    assert build.returncode == 0, build.stderr
    run = subprocess.run([str(exe)], capture_output=True, text=True, timeout=900)
    assert run.returncode == 0, run.stdout + run.stderr
    assert "zstd_frame_fuzz ok" in run.stdout and "ERROR" not in run.stderr


def test_property_random_structured_inputs_round_trip_product(host):
    """The same hypothesis property as the oracle's, on the product decoder's host build."""
    from hypothesis import given, settings, strategies as st

    frag = st.binary(min_size=1, max_size=40)

    @settings(max_examples=120, deadline=None)
    @given(frags=st.lists(frag, min_size=1, max_size=12), picks=st.lists(st.integers(0, 11), min_size=1, max_size=400),
           reps=st.lists(st.integers(1, 60), min_size=1, max_size=400), level=st.sampled_from([-3, 1, 2, 3, 5, 9, 13, 19, 22]))
    def run(frags, picks, reps, level):
        raw = b"".join(frags[p % len(frags)] * reps[i % len(reps)] for i, p in enumerate(picks))
        frame = O.libzstd_compress(raw, level)
        r, out = _decode(host, frame, len(raw))
        assert r == len(raw) and out == raw

    run()
