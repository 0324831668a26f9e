"""Model of csrc/lz4_block_cta.cuh, the one-CTA-per-block LZ4 decoder (test infrastructure, CPU only).

The CUDA decoder cannot run in the build container, so its ALGORITHM is restated here with the same
names and constants and checked against real LZ4 blocks (tests/test_lz4_cta_model.py):

  phase A  the input in super-steps of sixteen 256-byte windows: per window the successor code of
           every byte position (nx0), seven levels of pointer doubling over the in-window successors,
           the exit map, one thread chaining the true entry points through the exit maps, every window
           enumerating its real sequence starts as k-th successors, descriptors; the sequence that
           stops a chain (long length fields: "escape", or the block's last one) parsed serially;
  phase B  8 KiB tiles: literals and history bytes stored, every other match byte given a parent
           inside the tile (a repeating match points into its first period), pointer jumping a pair
           of bytes at a time -- both exit rules of l4_resolve --, root -> byte.

What the model does NOT restate: the shared-memory ring addressing, staging, the long-match list and
the split of B1 into fast paths -- they implement the same parent rule (l4_match_bytes) faster.
"""
import bisect

import numpy as np

WIN = 256          # kL4Win
STAGE = WIN + 32   # kL4Stage
WARPS = 16         # kL4Warps
LEVELS = 7         # kL4Levels
TILE = 8192        # kL4Tile
MAX_EXT = 4        # kL4MaxExt
LAST, ESC, BAD = 0xFFF1, 0xFFF2, 0xFFF3


def _literal_len(rd, p, limit, max_ext):
    """l4_literal_len: (code, lit, lit_pos); code 0 = fine"""
    lit = rd(p) >> 4
    q = p + 1
    if lit == 15:
        n = 0
        while True:
            if q >= limit:
                return BAD, 0, 0
            if n == max_ext:
                return ESC, 0, 0
            b = rd(q)
            q += 1
            lit += b
            n += 1
            if b != 255:
                break
    return 0, lit, q


def _nx0(inp, wb, p):
    """successor code of the candidate whose token sits at wb + p (phase A, step 1)"""
    in_size = len(inp)
    avail = in_size - wb
    if p >= avail:
        return BAD
    tok0 = inp[wb + p]
    if (tok0 >> 4) != 15 and (tok0 & 15) != 15:
        e = wb + p + 1 + (tok0 >> 4)
        return p + 3 + (tok0 >> 4) if e + 2 <= in_size else LAST if e == in_size else BAD
    lim = min(avail, STAGE) + wb   # the staged bytes end here
    r, lit, lit_pos = _literal_len(lambda i: inp[i], wb + p, lim, MAX_EXT)
    if r == BAD:
        r = ESC if lim < in_size else BAD   # ran off the STAGE, not the input
    if r:
        return r
    lit_end = lit_pos + lit
    if lit_end > in_size:
        return BAD
    if lit_end == in_size:
        return LAST
    if in_size - lit_end < 2:
        return BAD
    q = lit_end + 2
    if (tok0 & 15) == 15:
        n, b = 0, 255
        while b == 255:
            if q >= in_size:
                return BAD
            if n == MAX_EXT or q >= lim:
                return ESC
            b = inp[q]
            q += 1
            n += 1
    rel = q - wb
    return rel if rel < 0xFFF0 else ESC


def parse(inp: bytes, out_cap: int, stats=None):
    """l4_parse: (nseq or negative error, descriptors [(out_pos, lit_pos, lit, off)], total_out)"""
    in_size = len(inp)
    desc = []
    pos = out = 0
    done = False
    steps = 0
    while not done:
        B = pos
        steps += 1
        nx0, f, ex = {}, {}, {}
        for w in range(WARPS):
            wb = B + w * WIN
            if wb >= in_size:
                continue
            code = [_nx0(inp, wb, p) for p in range(WIN)]
            lv = [[c if c < WIN else p for p, c in enumerate(code)]]
            for _ in range(1, LEVELS):
                a = lv[-1]
                lv.append([a[a[p]] for p in range(WIN)])
            a = lv[-1]
            nx0[w], f[w] = code, lv
            ex[w] = [code[a[a[p]]] for p in range(WIN)]   # 2^6 + 2^6 steps: the last start inside the window
        # 2. one thread chains the true entry points
        entry = {}
        e = w = 0
        stop_code = stop_pos = 0
        nxt = B
        while w < WARPS and B + w * WIN < in_size:
            entry[w] = e
            x = ex[w][e]
            if x >= 0xFFF0:
                a = f[w][LEVELS - 1]
                stop_code, stop_pos = x, B + w * WIN + a[a[e]]
                break
            e, w = x & 255, w + (x >> 8)
            nxt = B + w * WIN + e
        # 3. + 4. every window enumerates its real starts (k-th successors), parses them, descriptors
        for w in sorted(entry):
            wb = B + w * WIN
            prev = None
            for k in range(96):
                p = entry[w]
                for L in range(LEVELS):
                    if k & (1 << L):
                        p = f[w][L][p]
                code = nx0[w][p]
                valid = (k == 0 or p != prev) and code < 0xFFF0
                prev = p
                if not valid:
                    continue
                tok = inp[wb + p]
                lit, qq = tok >> 4, wb + p + 1
                if lit == 15:
                    while True:
                        b = inp[qq]
                        qq += 1
                        lit += b
                        if b != 255:
                            break
                mq = qq + lit
                off = inp[mq] | (inp[mq + 1] << 8)
                mq += 2
                ml = (tok & 15) + 4
                if (tok & 15) == 15:
                    while True:
                        b = inp[mq]
                        mq += 1
                        ml += b
                        if b != 255:
                            break
                desc.append((out, qq, lit, off))
                out += lit + ml
        if out > out_cap:
            return -4, desc, out
        # the sequence that stopped the chain
        if stop_code:
            p = stop_pos
            if stop_code == BAD:
                return -1, desc, out
            r, lit, lit_pos = _literal_len(lambda i: inp[i], p, in_size, 1 << 62)
            if r or lit > in_size - lit_pos:
                return -2, desc, out
            if lit_pos + lit == in_size:
                if lit > out_cap - out:
                    return -2, desc, out
                desc.append((out, lit_pos, lit, 0))
                out += lit
                done = True
            elif in_size - (lit_pos + lit) < 2:
                return -3, desc, out
            else:
                qq = lit_pos + lit + 2
                off = inp[qq - 2] | (inp[qq - 1] << 8)
                ml = inp[p] & 15
                if ml == 15:
                    b = 255
                    while b == 255:
                        if qq >= in_size:
                            return -1, desc, out
                        b = inp[qq]
                        qq += 1
                        ml += b
                        if ml > out_cap:
                            return -4, desc, out
                ml += 4
                if lit > out_cap - out or ml > out_cap - out - lit:
                    return -4, desc, out
                desc.append((out, lit_pos, lit, off if off else 0x10000))
                out += lit + ml
                nxt = qq
                if qq >= in_size:
                    return -1, desc, out
        elif nxt >= in_size:
            return -1, desc, out
        pos = nxt
    if stats is not None:
        stats["super_steps"] = steps
        stats["sequences"] = len(desc)
    return len(desc), desc, out


def resolve(P, early_root):
    """l4_resolve: pointer jumping over the parents of one tile, a pair of bytes at a time, two hops per
    round; a round reads P[] as it was when the round began.  Returns the rounds it took."""
    n = len(P)
    x = np.arange(0, n, 2)
    pp = np.stack([P[x], P[x + 1]], axis=1)
    act = ~((pp[:, 0] == x) & (pp[:, 1] == x + 1))
    rounds = 0
    while True:
        rounds += 1
        snap = P.copy()
        changed = np.zeros(len(x), dtype=bool)
        idx = np.nonzero(act)[0]
        h1 = np.stack([snap[pp[idx, 0]], snap[pp[idx, 1]]], axis=1)
        h2 = np.stack([snap[h1[:, 0]], snap[h1[:, 1]]], axis=1)
        if early_root:
            same1 = (h1 == pp[idx]).all(axis=1)
            act[idx[same1]] = False
            go = idx[~same1]
            pp[go] = h2[~same1]
            changed[go] = True
            act[go[(h2[~same1] == h1[~same1]).all(axis=1)]] = False
        else:
            mv = (h2 != pp[idx]).any(axis=1)
            pp[idx[mv]] = h2[mv]
            changed[idx[mv]] = True
            act[idx[~mv]] = False
        P[x[changed]] = pp[changed, 0]
        P[x[changed] + 1] = pp[changed, 1]
        if (not act.any()) if early_root else (not changed.any()):
            return rounds


def copy(inp: bytes, desc, total: int, ga: int = 0, early_root: bool = False, stats=None):
    """l4_copy: descriptors -> bytes, tile by tile (v-space: v = out_pos + ga).  Returns (status, bytes)."""
    in_size = len(inp)
    out = np.zeros(total + ga, dtype=np.uint8)   # index = v
    src = np.frombuffer(inp, dtype=np.uint8)
    ends = [d[0] for d in desc[1:]] + [total]
    vend = total + ga
    rounds = tiles = 0
    for tlo in range(0, vend, TILE):
        thi = min(tlo + TILE, vend)
        lo_v = ga if tlo == 0 else tlo
        P = np.arange(TILE, dtype=np.int64)
        k = bisect.bisect_right(ends, lo_v - ga)              # the first sequence that ends inside or behind the tile's start
        while k < len(desc) and desc[k][0] + ga < thi:
            o, lit_pos, lit, off = desc[k]
            o += ga
            oe = ends[k] + ga
            k += 1
            if oe < o or lit_pos > in_size or lit > in_size - lit_pos or lit > oe - o:
                return -2, b""
            m, ml = o + lit, oe - (o + lit)
            a, b = max(o, lo_v), min(m, thi)
            if a < b:
                out[a:b] = src[lit_pos + (a - o):lit_pos + (b - o)]
            if ml == 0:
                continue
            if off == 0 or off > 0xFFFF or off > m - ga:
                return -4, b""
            a, b = max(m, lo_v), min(oe, thi)
            if a >= b:
                continue
            v = np.arange(a, b)
            d = v - m
            sp = (m - off) + (d % off if off < ml else d)   # l4_match_bytes: a repeating match stays in its first period
            hist = sp < tlo
            out[v[hist]] = out[sp[hist]]                      # before the tile: final bytes
            P[v[~hist] - tlo] = sp[~hist] - tlo
        assert (P <= np.arange(TILE)).all()
        rounds += resolve(P, early_root)
        tiles += 1
        n = thi - tlo
        out[tlo:thi] = out[tlo + P[:n]]                       # root -> byte (roots are final since B1)
    if stats is not None:
        stats["tiles"] = tiles
        stats["rounds"] = rounds
    return total, out[ga:].tobytes()


def decode(inp: bytes, out_cap: int, ga: int = 0, early_root: bool = False, stats=None):
    nseq, desc, total = parse(inp, out_cap, stats)
    if nseq < 0:
        return nseq, b""
    return copy(inp, desc, total, ga, early_root, stats)
