#!/usr/bin/env python3
"""Small driver for compute-sanitizer: every kernel variant, both modes, ragged / unaligned
inputs, the fused exchange at world 1, the block stream, the LZ4 and Zstd decoders, on small inputs."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import libflagstats_b200 as fs
from libflagstats_b200 import sharded, synth
from oracle import oracle as O

lib = fs.lib()
for variant in range(9):
    if lib.FLAGSTAT_cuda_set_variant(variant) < 0:
        continue  # an A/B variant that is not compiled into the product library
    for n, off in ((0, 0), (5, 1), (16384 * 8 * 3 + 77, 3), (1_300_003, 0)):
        d = synth.uniform_device(n + off, 0, 7, 0x0FFF)[off:]
        want = O.flagstat_simd(O.synth_uniform(off, n, 7, 0x0FFF))
        assert fs.flagstat_u64(d).tolist() == want.tolist(), (variant, n, off)
    a = O.synth_uniform(0, 200_001, 3, 0xFFFF)
    assert fs.pospopcnt_u16(a).tolist() == O.pospopcnt(a).tolist()
lib.FLAGSTAT_cuda_set_variant(0)
# the dynamically scheduled kernel forced onto short columns (both chunk sizes), back to back
for cg in (1, 2):
    lib.FLAGSTAT_cuda_set_dynamic(1, cg)
    for n, off in ((4096 * 3 + 5, 1), (1_300_003, 0), (4096 * 2368 + 11, 3)):
        d = synth.uniform_device(n + off, 0, 7, 0x0FFF)[off:]
        want = O.flagstat_simd(O.synth_uniform(off, n, 7, 0x0FFF))
        for _ in range(3):
            assert fs.flagstat_u64(d).tolist() == want.tolist(), ("dynamic", cg, n, off)
lib.FLAGSTAT_cuda_set_dynamic(0, 1)
x = sharded.FusedExchange()
d = synth.hiseqx_device(700_001, 0, 1, 3000)
out = x.flagstat(d)
torch.cuda.synchronize()
assert out.cpu().numpy().view(np.uint64).tolist() == O.flagstat_simd(O.synth_hiseqx(0, 700_001, 1, 3000)).tolist()
# samtools mode (exact n_pair_all) through the host, device and exchange entries
u = O.synth_uniform(0, 400_003, 5, 0xFFFF)
assert fs.samtools_stats(u).tolist() == O.samtools_loop(u).tolist()
du = synth.uniform_device(400_003 + 3, 0, 5, 0xFFFF)[3:]
want = O.flagstat_simd(O.synth_uniform(3, 400_003, 5, 0xFFFF))
st = O.samtools_loop(O.synth_uniform(3, 400_003, 5, 0xFFFF))
want[0], want[16] = np.uint64(st[2, 0]), np.uint64(st[2, 1])
assert fs.flagstat_samtools_u64(du).tolist() == want.tolist()
assert x.flagstat(du, samtools=True).cpu().numpy().view(np.uint64).tolist() == want.tolist()
# overlapped steps (programmatic dependent launch) on a side stream
x.set_overlap(True)
side = torch.cuda.Stream()
torch.cuda.synchronize()
with torch.cuda.stream(side):
    acc = torch.zeros(32, dtype=torch.int64, device="cuda")
    for _ in range(6):
        x.flagstat(d, out=acc, accumulate=True, stream=side)
        x.flagstat(du, out=acc, accumulate=True, stream=side)
side.synchronize()
w6 = 6 * (O.flagstat_simd(O.synth_hiseqx(0, 700_001, 1, 3000)) + O.flagstat_simd(O.synth_uniform(3, 400_003, 5, 0xFFFF)))
assert acc.cpu().numpy().view(np.uint64).tolist() == w6.tolist()
x.status()
x.close()
a = O.synth_uniform(0, 300_000, 9, 0x0FFF)
for mode in (0, 1):
    with fs.BlockStream(0, 40_000, 2, mode=mode, coalesce=3) as bs:
        for lo in range(0, a.size, 40_000):
            bs.push(a[lo:lo + 40_000])
        assert bs.finish().tolist() == O.flagstat_simd(a).tolist()
# LZ4 block decoders (both), aligned and packed (unaligned) outputs, every copy path
from libflagstats_b200 import blockfile  # noqa: E402
from tests.test_blockfile import _handmade_chain_block  # noqa: E402

rng = np.random.default_rng(11)
cats = np.array([99, 147, 83, 163, 97, 145, 73, 137, 2113, 77], np.uint16)
cols = [np.repeat(cats[rng.integers(0, 10, 3000)], rng.geometric(1 / 8, 3000)),
        cats[rng.integers(0, 10, 20_000)],
        np.concatenate([O.synth_uniform(0, 12_000, 9, 0x0FFF)] * 2),
        np.repeat(rng.integers(0, 4096, 40).astype(np.uint16), rng.integers(1, 3000, 40)),
        O.synth_hiseqx(0, 50_001, 2, 1000)]
blocks = [O.liblz4_compress(c.tobytes()) for c in cols] + [O.lz4_compress(c.tobytes()) for c in cols]
raws = [c.tobytes() for c in cols] * 2
hb, hr = _handmade_chain_block()
blocks.append(hb)
raws.append(hr)
# malformed blocks too (byte flips, truncations, runs of 0xFF length bytes): whatever the verdict, no
# access outside the buffers and no use of uninitialised shared memory
bad, bad_sizes = [], []
for comp, raw in zip(blocks[:5], raws[:5]):
    for k in range(6):
        b = bytearray(comp)
        if k % 3 == 0:
            b[int(rng.integers(0, len(b)))] = int(rng.integers(0, 256))
        elif k % 3 == 1:
            b = b[: int(rng.integers(1, len(b)))]
        else:
            i = int(rng.integers(0, len(b)))
            b[i:i + 1] = bytes([0xFF] * int(rng.integers(1, 7)))
        bad.append(bytes(b))
        bad_sizes.append(len(raw))
for v in (2, 1, 0):
    lib.FLAGSTAT_cuda_set_lz4_variant(v)
    out, status = blockfile.lz4_decode(blocks, [len(r) for r in raws])
    assert status == [len(r) for r in raws], (v, status)
    assert all(o == r for o, r in zip(out, raws)), v
    blockfile.lz4_decode(bad, bad_sizes)
    blob = O.write_lz4_container(cols[0], block_bytes=20_002)
    f, n = blockfile.flagstat_container(blob, blockfile.LZ4)
    assert n == cols[0].size and f.tolist() == O.numpy_flagstat(cols[0]).tolist(), v
lib.FLAGSTAT_cuda_set_lz4_variant(2)
# Zstd frames, both decoders (two-stage default; one thread per frame): well-formed frames of several shapes
# (far offsets, noise = raw blocks, tiny), corrupted frames, a container
if O.libzstd() is not None:
    noise = rng.integers(0, 256, 70_001, dtype=np.uint8).tobytes()
    zraws = [c.tobytes() for c in cols[:3]] + [noise * 3 + noise[:999], bytes(rng.integers(0, 3, 50_001, dtype=np.uint8)), b"xyz"]
    zframes = [O.libzstd_compress(r, lvl) for r, lvl in zip(zraws, (1, 3, 19, 3, 1, 1))]
    zbad = []
    for f in zframes[:4]:
        for k in range(6):
            b = bytearray(f)
            if k % 2 == 0:
                b[int(rng.integers(4, len(b)))] ^= 1 << int(rng.integers(0, 8))
            else:
                b = b[: int(rng.integers(1, len(b)))]
            zbad.append(bytes(b))
    for variant in ("1", "0"):
        os.environ["FLAGSTAT_CUDA_ZSTD_VARIANT"] = variant
        out, status = blockfile.zstd_decode(zframes, [len(r) for r in zraws])
        assert status == [len(r) for r in zraws], (variant, status)
        assert all(o == r for o, r in zip(out, zraws)), variant
        blockfile.zstd_decode(zbad, [len(zraws[i // 6]) for i in range(len(zbad))])
        blob = O.write_zstd_container(cols[0], 1)
        f, n = blockfile.flagstat_container(blob, blockfile.ZSTD)
        assert n == cols[0].size and f.tolist() == O.numpy_flagstat(cols[0]).tolist(), variant
    os.environ["FLAGSTAT_CUDA_ZSTD_VARIANT"] = "1"
# pageable host arrays: threaded staging
a = O.synth_hiseqx(0, 6_000_001, 4, 5000)
assert fs.flagstat_u64(a).tolist() == O.flagstat_simd(a).tolist()
print("sanitize driver ok")
