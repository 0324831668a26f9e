#!/usr/bin/env python3
"""Streamed end-to-end (BASELINE configs[4]): 1,024,000-byte blocks from a pinned ring.
Sweeps transport (DMA / zero-copy), blocks per launch and ring depth; the loop runs in C
(FLAGSTAT_cuda_stream_selftime), so no Python overhead is inside the timed region.

    python tools/stream_bench.py [n_blocks] > gpurun_out/stream.jsonl
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import libflagstats_b200 as fs
from libflagstats_b200 import synth


def pcie_probe():
    n = 256 << 20
    h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(4):
        d.copy_(h, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    return 4 * n / (e0.elapsed_time(e1) * 1e-3) / 1e9


def main():
    n_blocks = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
    probe = pcie_probe()
    print(json.dumps({"case": "pinned cudaMemcpyAsync H2D probe", "gbs": probe}), flush=True)
    col = synth.hiseqx_device(64 * fs.BLOCK_RECORDS * 4, 0, 3, 0).cpu().numpy().view(np.uint16)
    for mode in (0, 1):
        for coalesce in (1, 2, 4, 8, 16):
            for slots in (2, 3, 4):
                with fs.BlockStream(0, fs.BLOCK_RECORDS, slots, mode=mode, coalesce=coalesce) as bs:
                    for i in range(slots * coalesce):
                        s = bs.acquire()
                        s[:] = col[i * fs.BLOCK_RECORDS:(i + 1) * fs.BLOCK_RECORDS]
                        bs.submit(fs.BLOCK_RECORDS)
                    bs.finish()
                    bs.selftime(200)
                    best = None
                    for _ in range(3):
                        f, sec = bs.selftime(n_blocks)
                        best = sec if best is None else min(best, sec)
                    gbs = n_blocks * 1.024e-3 / best
                    print(json.dumps({"case": "stream 1,024,000-B blocks", "mode": ["dma", "zerocopy"][mode],
                                      "blocks_per_launch": coalesce, "ring_groups": slots, "n_blocks": n_blocks,
                                      "gbs": round(gbs, 2), "grec_s": round(gbs / 2, 2),
                                      "frac_of_pcie_probe": round(gbs / probe, 3),
                                      "us_per_block": round(best / n_blocks * 1e6, 2),
                                      "records": int(f[9] + f[25])}), flush=True)


if __name__ == "__main__":
    main()
