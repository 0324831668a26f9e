#!/usr/bin/env python3
"""CPU study behind DESIGN.md section 12 item 3 (Zstd decoder, second version): for the bench
columns, compressed by the real libzstd in the reference's 1,024,000-byte blocks, how many
Zstd blocks and sequences does a frame hold, and how are the literals coded?  Walks the frame
structure only (RFC 8878 3.1.1: block headers, literals section header, sequence count).

    python tools/zstd_sequence_study.py [n_blocks] [levels...] > profiles/<tag>_zstd_sequence_study.jsonl
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from oracle import oracle as O  # noqa: E402
from file_bench import runs_column  # noqa: E402


def walk(frame: bytes):
    """[(block type, regenerated literals, literal coding, sequences)] of one frame."""
    fhd = frame[4]
    single, dict_flag, fcs_flag = (fhd >> 5) & 1, fhd & 3, fhd >> 6
    ip = 5 + (0 if single else 1) + (0, 1, 2, 4)[dict_flag]
    ip += (1 if single else 0) if fcs_flag == 0 else (2, 4, 8)[fcs_flag - 1]
    out = []
    while True:
        bh = int.from_bytes(frame[ip:ip + 3], "little"); ip += 3
        last, btype, size = bh & 1, (bh >> 1) & 3, bh >> 3
        if btype == 2:
            p = frame[ip:ip + size]
            ltype, sf = p[0] & 3, (p[0] >> 2) & 3
            if ltype < 2:
                hdr = 1 if sf in (0, 2) else 2 if sf == 1 else 3
                regen = int.from_bytes(p[:hdr], "little") >> (3 if hdr == 1 else 4)
                comp = regen if ltype == 0 else 1
            else:
                hdr = 3 if sf < 2 else 4 if sf == 2 else 5
                bits = 10 if sf < 2 else 14 if sf == 2 else 18
                v = int.from_bytes(p[:hdr], "little")
                regen, comp = (v >> 4) & ((1 << bits) - 1), (v >> (4 + bits)) & ((1 << bits) - 1)
            q = p[hdr + comp:]
            nseq = q[0] if q[0] < 128 else ((q[0] - 128) << 8) + q[1] if q[0] < 255 else q[1] + (q[2] << 8) + 0x7F00
            out.append((("raw", "rle", "huffman", "treeless")[ltype], regen, nseq))
            ip += size
        else:
            out.append((("rawblock", "rleblock")[btype], 0, 0))
            ip += size if btype == 0 else 1
        if last:
            return out


def main():
    n_blocks = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    levels = [int(x) for x in sys.argv[2:]] or [1, 3, 19]
    n = n_blocks * 512_000
    rng = np.random.default_rng(3)
    cats = np.array([99, 147, 83, 163, 97, 145, 73, 137, 133, 69, 77, 141, 2113, 2177], np.uint16)
    p = np.array([195, 195, 195, 195, 8.4, 8.4, 1, 1, 1, 1, 8.5, 8.5, 1.3, 1.3])
    cols = {
        "hiseqx_generator (quasi-periodic)": O.synth_hiseqx(0, n),
        "hiseqx_categories_runs_mean8": runs_column(n),
        "hiseqx_categories_iid": rng.choice(cats, size=n, p=p / p.sum()).astype(np.uint16),
    }
    for name, col in cols.items():
        raw = col.tobytes()
        for level in levels:
            comp = nblk = nseq = nlit = 0
            kinds = {}
            frames = 0
            for lo in range(0, len(raw), O.REF_BLOCK_BYTES):
                frame = O.libzstd_compress(raw[lo:lo + O.REF_BLOCK_BYTES], level)
                comp += len(frame)
                frames += 1
                for kind, regen, ns in walk(frame):
                    nblk += 1
                    nseq += ns
                    nlit += regen
                    kinds[kind] = kinds.get(kind, 0) + 1
            print(json.dumps({"column": name, "level": level, "ratio": round(len(raw) / comp, 2),
                              "zstd_blocks_per_frame": round(nblk / frames, 1), "sequences_per_frame": int(nseq / frames),
                              "bytes_per_sequence": round(len(raw) / max(1, nseq), 1),
                              "literal_bytes_per_frame": int(nlit / frames), "literal_coding_of_blocks": kinds}), flush=True)


if __name__ == "__main__":
    main()
