#!/usr/bin/env python3
"""CPU study behind DESIGN.md section 12 item 2 (segment-parallel LZ4 decode): for the three
column shapes of tools/file_bench.py, compressed by the real liblz4 in the reference's
1,024,000-byte blocks, how many sequences does a block hold, how far back do matches reach,
and how many matches of a 64 KiB output segment read bytes that another segment produces
(= would have to wait on that segment's progress in a look-back scheme)?

    python tools/lz4_segment_study.py [n_blocks] > profiles/<tag>_lz4_segment_study.jsonl
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

SEG = 64 << 10


def sequences(block: bytes):
    """(literal length, match length, offset, output position of the match) per sequence."""
    ip, op, n = 0, 0, len(block)
    out = []
    while ip < n:
        tok = block[ip]; ip += 1
        lit = tok >> 4
        if lit == 15:
            while True:
                b = block[ip]; ip += 1
                lit += b
                if b != 255:
                    break
        ip += lit
        op += lit
        if ip >= n:
            break
        off = block[ip] | (block[ip + 1] << 8); ip += 2
        ml = (tok & 15) + 4
        if (tok & 15) == 15:
            while True:
                b = block[ip]; ip += 1
                ml += b
                if b != 255:
                    break
        out.append((lit, ml, off, op))
        op += ml
    return out, op


def main():
    n_blocks = int(sys.argv[1]) if len(sys.argv) > 1 else 6
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
        nseq, cross, waits_far, offs, mls, lits, comp = [], [], [], [], [], [], 0
        for lo in range(0, len(raw), O.REF_BLOCK_BYTES):
            blk = O.liblz4_compress(raw[lo:lo + O.REF_BLOCK_BYTES])
            comp += len(blk)
            seqs, produced = sequences(blk)
            assert produced == min(O.REF_BLOCK_BYTES, len(raw) - lo)
            nseq.append(len(seqs))
            c = f = 0
            for lit, ml, off, op in seqs:
                seg_start = op // SEG * SEG
                if op - off < seg_start:          # source starts in an earlier segment
                    c += 1
                    if op - off < seg_start - SEG:  # ... and not even in the previous one
                        f += 1
                offs.append(off); mls.append(ml); lits.append(lit)
            cross.append(c); waits_far.append(f)
        offs, mls, lits = np.array(offs), np.array(mls), np.array(lits)
        print(json.dumps({
            "column": name, "blocks": len(nseq), "ratio": round(len(raw) / comp, 2),
            "sequences_per_block": int(np.mean(nseq)),
            "match_length": {"mean": round(float(mls.mean()), 1), "p50": int(np.percentile(mls, 50)), "p99": int(np.percentile(mls, 99))},
            "literal_length": {"mean": round(float(lits.mean()), 2), "p99": int(np.percentile(lits, 99))},
            "offset": {"p50": int(np.percentile(offs, 50)), "p90": int(np.percentile(offs, 90)), "p99": int(np.percentile(offs, 99)),
                       "max": int(offs.max())},
            "segments_per_block": -(-O.REF_BLOCK_BYTES // SEG),
            "matches_reading_an_earlier_segment_per_block": round(float(np.mean(cross)), 1),
            "share_of_matches": round(float(np.sum(cross)) / max(1, int(np.sum(nseq))), 5),
            "of_those_beyond_the_previous_segment_per_block": round(float(np.mean(waits_far)), 1),
        }), flush=True)


if __name__ == "__main__":
    main()
