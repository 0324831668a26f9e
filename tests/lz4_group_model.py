"""Lock-step model of csrc/lz4_block_group.cuh (test infrastructure, CPU only).

The CUDA decoder cannot run in the build container, so its ALGORITHM -- window
staging, the len[] chase, the warp scan, dependency rounds between matches of one
group, ring addressing with the global-alignment shift, partial flushes -- is
restated here with the 32 lanes as explicit loops and the same variable names,
and checked against real LZ4 blocks (tests/test_lz4_oracle.py).  Lanes inside a
phase are executed in a SCRAMBLED order, so a missing dependency or a missing
__syncwarp() in the design shows up as a wrong byte.
"""
import random

WIN = 8192    # kGrpRing: bytes of recent output mirrored in shared memory
RING = WIN
GW = 256      # kGrpWin
OUTCAP = 2048  # kGrpOutCap: output bytes of one group (entries of P[])


class Out:
    def __init__(self, cap, ga):
        self.g = bytearray(b"\xEE" * cap)     # "global memory"
        self.ring = bytearray(WIN)
        self.ga = ga
        self.flushed = 0

    def ridx(self, pos):
        return (pos + self.ga) & (WIN - 1)

    def put(self, pos, v):
        self.g[pos] = v
        self.ring[self.ridx(pos)] = v

    def flush(self, upto, exact):
        hi = upto
        if not exact:
            a = (self.ga + upto) & ~15
            hi = a - self.ga if a > self.ga else 0
        if hi <= self.flushed:
            return
        lo = self.flushed
        mis = (self.ga + lo) & 15
        if mis:
            hb = min(16 - mis, hi - lo)
            for lane in range(hb):
                self.g[lo + lane] = self.ring[self.ridx(lo + lane)]
            lo += hb
        nv = (hi - lo) >> 4
        for v in range(nv):
            p = lo + (v << 4)
            assert (self.ga + p) % 16 == 0 and self.ridx(p) % 16 == 0 and self.ridx(p) + 16 <= WIN
            self.g[p:p + 16] = self.ring[self.ridx(p):self.ridx(p) + 16]
        lo += nv << 4
        assert hi - lo < 16 and (exact or hi == lo)
        for lane in range(hi - lo):
            self.g[lo + lane] = self.ring[self.ridx(lo + lane)]
        self.flushed = hi


def _slow_sequence(inp, o, out_cap, ip, op):
    """lz4g_slow_sequence: returns (rc, ip, op)."""
    n = len(inp)
    token = inp[ip]; ip += 1
    lit = token >> 4
    if lit == 15:
        while True:
            if ip >= n:
                return -1, ip, op
            b = inp[ip]; ip += 1
            lit += b
            if b != 255:
                break
    if lit > n - ip or lit > out_cap - op:
        return -2, ip, op
    lit_at = ip
    ip += lit
    last = ip >= n
    offset = ml = 0
    if not last:
        if n - ip < 2:
            return -3, ip, op
        offset = inp[ip] | (inp[ip + 1] << 8)
        ip += 2
        ml = token & 15
        if ml == 15:
            while True:
                if ip >= n:
                    return -1, ip, op
                b = inp[ip]; ip += 1
                ml += b
                if b != 255:
                    break
        ml += 4
        if offset == 0 or offset > op + lit or ml > out_cap - op - lit:
            return -4, ip, op
    for i in range(lit):
        o.put(op + i, inp[lit_at + i])
    op += lit
    if last:
        return 1, ip, op
    from_ring = offset + ml <= WIN
    base = op - offset
    # warp_match: every byte's source is base + (i % offset); all reads of a step precede its writes only
    # within one 32*unroll step of non-overlapping sources -- equivalent to the sequential definition
    for i in range(ml):
        s = base + (i % offset)
        v = o.ring[o.ridx(s)] if from_ring else o.g[s]
        assert s < op  # produced earlier: flushed before this call, or this sequence's literals (put)
        o.put(op + i, v)
    op += ml
    return (0 if ip < n else 1), ip, op


def decode(inp, out_cap, ga=0, seed=0, stats=None):
    """Returns (status, bytes).  status = decoded size or a negative code."""
    rnd = random.Random(seed)
    inp = bytes(inp)
    n = len(inp)
    o = Out(out_cap + 64, ga)
    ip = op = 0
    if n == 0:
        return 0, b""
    win = bytearray(GW)
    lent = bytearray(GW)
    while True:
        avail = n - ip
        K = consumed = 0
        t0 = inp[ip]
        posv = []
        simple = (t0 >> 4) != 15
        if simple and (t0 & 15) == 15:
            e = ip + 3 + (t0 >> 4)
            simple = e < n and inp[e] != 255
        if simple:
            for p in range(GW):
                win[p] = inp[ip + p] if p < avail else 0
            for p in range(GW):
                b = win[p]
                lit, mln = b >> 4, b & 15
                ln = 3 + lit if lit != 15 else 0
                if mln == 15 and ln:
                    ln += 1
                    if p + ln >= GW or win[p + ln - 1] == 255:
                        ln = 0
                if p + ln >= GW or p + ln >= avail:   # successor positions are one byte (nx[] tables)
                    ln = 0
                lent[p] = ln
            p = 0
            while len(posv) < 32 and p < GW:
                ln = lent[p]
                if ln == 0:
                    break
                posv.append(p)
                p += ln
            K = len(posv)
            posv.append(p)              # posv[K] = end of the last sequence
        if K == 0:
            o.flush(op, True)
            assert o.flushed == op
            rc, ip, op = _slow_sequence(inp, o, out_cap, ip, op)
            o.flushed = op
            if stats is not None:
                stats["slow"] = stats.get("slow", 0) + 1
            if rc < 0:
                return rc, b""
            if rc == 1:
                break
            continue
        lanes = list(range(32))
        lit = [0] * 32; ml = [0] * 32; off = [0] * 32; pk = [0] * 32; ext = [False] * 32
        for k in range(32):
            pk[k] = posv[k] if k < K else 0
            tok = win[pk[k]]
            off[k] = win[pk[k] + 1 + (tok >> 4)] | (win[pk[k] + 2 + (tok >> 4)] << 8)
            if k < K:
                lit[k] = tok >> 4
                ext[k] = (tok & 15) == 15
                ml[k] = (tok & 15) + 4 + (win[pk[k] + 3 + (tok >> 4)] if ext[k] else 0)
        incl, acc = [0] * 32, 0
        for k in range(32):
            acc += lit[k] + ml[k]
            incl[k] = acc
        if incl[K - 1] > OUTCAP:        # P[] holds OUTCAP bytes: cut the group (one sequence is <= 287)
            K = max(k + 1 for k in range(K) if incl[k] <= OUTCAP)
            for k in range(K, 32):
                lit[k] = ml[k] = 0
            incl = [incl[min(k, K - 1)] for k in range(32)]
        consumed = posv[K]
        total = incl[31]
        if total > out_cap - op:
            return -4, b""
        o_k = [op + incl[k] - lit[k] - ml[k] for k in range(32)]
        m_k = [o_k[k] + lit[k] for k in range(32)]
        if any(k < K and (off[k] == 0 or off[k] > m_k[k]) for k in range(32)):
            return -4, b""
        rnd.shuffle(lanes)
        for k in lanes:
            for i in range(lit[k]):
                o.ring[o.ridx(o_k[k] + i)] = win[pk[k] + 1 + i]
        # 4b. matches: per-byte parent pointers P[] (relative to op), then pointer jumping
        P = list(range(total))          # literal bytes are their own root
        order = list(range(32))
        rnd.shuffle(order)
        for k in order:
            if k >= K:
                continue
            src = m_k[k] - off[k]
            for i in range(ml[k]):
                s_ = src + (i % off[k] if off[k] < ml[k] else i)
                rel = m_k[k] + i - op
                if s_ < op:             # produced before this group: copy now, byte becomes a root
                    if (op + total) - s_ <= RING:
                        v = o.ring[o.ridx(s_)]
                    else:
                        assert s_ < o.flushed
                        v = o.g[s_]
                    o.ring[o.ridx(op + rel)] = v
                    P[rel] = rel
                else:
                    assert s_ - op < rel
                    P[rel] = s_ - op
        rounds = 0
        while True:
            rounds += 1
            changed = False
            bs = list(range(total))
            rnd.shuffle(bs)             # in-place, any order: every value read is an ancestor
            for b in bs:
                p_ = P[b]
                if p_ != b:
                    q = P[p_]
                    if q != p_:
                        P[b] = q
                        changed = True
            if not changed:
                break
        bs = list(range(total))
        rnd.shuffle(bs)
        for b in bs:
            r = P[b]
            if r != b:
                assert P[r] == r
                o.ring[o.ridx(op + b)] = o.ring[o.ridx(op + r)]
        if stats is not None:
            stats["groups"] = stats.get("groups", 0) + 1
            stats["seqs"] = stats.get("seqs", 0) + K
            stats["rounds"] = stats.get("rounds", 0) + rounds
        ip += consumed
        op += total
        assert ip < n
        o.flush(op, False)
    o.flush(op, True)
    return op, bytes(o.g[:op])
