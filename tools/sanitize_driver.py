#!/usr/bin/env python3
"""Small driver for compute-sanitizer: every kernel variant, both modes, ragged / unaligned
inputs, the fused exchange at world 1 and the block stream, on small inputs."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import libflagstats_b200 as fs
from libflagstats_b200 import sharded, synth
from oracle import oracle as O

lib = fs.lib()
for variant in range(8):
    lib.FLAGSTAT_cuda_set_variant(variant)
    for n, off in ((0, 0), (5, 1), (16384 * 8 * 3 + 77, 3), (1_300_003, 0)):
        d = synth.uniform_device(n + off, 0, 7, 0x0FFF)[off:]
        want = O.flagstat_simd(O.synth_uniform(off, n, 7, 0x0FFF))
        assert fs.flagstat_u64(d).tolist() == want.tolist(), (variant, n, off)
    a = O.synth_uniform(0, 200_001, 3, 0xFFFF)
    assert fs.pospopcnt_u16(a).tolist() == O.pospopcnt(a).tolist()
lib.FLAGSTAT_cuda_set_variant(0)
x = sharded.FusedExchange()
d = synth.hiseqx_device(700_001, 0, 1, 3000)
out = x.flagstat(d)
torch.cuda.synchronize()
assert out.cpu().numpy().view(np.uint64).tolist() == O.flagstat_simd(O.synth_hiseqx(0, 700_001, 1, 3000)).tolist()
x.close()
a = O.synth_uniform(0, 300_000, 9, 0x0FFF)
for mode in (0, 1):
    with fs.BlockStream(0, 40_000, 2, mode=mode, coalesce=3) as bs:
        for lo in range(0, a.size, 40_000):
            bs.push(a[lo:lo + 40_000])
        assert bs.finish().tolist() == O.flagstat_simd(a).tolist()
print("sanitize driver ok")
