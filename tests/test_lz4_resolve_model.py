"""Model of the pointer-jumping rounds of the LZ4 CTA decoder's copy phase
(libflagstats_b200/csrc/lz4_block_cta.cuh, l4_resolve): every byte of a tile has a parent P[x] <= x
(the byte it copies; roots -- literals, history bytes -- are their own parent); a round reads P[] as it
was when the round began (the barrier), two hops per pair of bytes, then the owners store.

Two rules for when a PAIR of bytes stops taking part are modelled:
  * "verify":     a pair stops when a round did not move its parents (one whole extra round per pair),
  * "early root": a pair stops as soon as it knows that its parents are roots -- after the first hop
                  if that did not move them, after the second if the first hop's result did not move
                  (-DFSB_L4_EARLY_ROOT=1).
Both must end with P[x] = root(x) for every byte; the second must never need more hops.
CPU only: this checks the protocol, the GPU parity tests (tests/test_blockfile.py) check the kernel.
"""
import numpy as np
import pytest


def make_forest(rng, n, kind):
    """parents of a tile of n bytes laid out the way LZ4 matches produce them"""
    P = np.arange(n, dtype=np.int64)
    x = 0
    while x < n:
        if kind == "iid":        # short matches, random offsets: chains of depth ~ n / mean offset
            lit = int(rng.integers(0, 2)) if x else 4
            ml = int(rng.integers(4, 12))
            off = int(rng.integers(1, 200))
        elif kind == "runs":     # literals + a long repeating match with offset 2
            lit, ml, off = 2, int(rng.integers(4, 120)), 2
        elif kind == "deep":     # offset-1..3 plain copies everywhere: the deepest chains a tile can hold
            lit = 1 if x == 0 else 0
            ml, off = int(rng.integers(1, 4)), int(rng.integers(1, 4))
        else:                    # mixed, odd alignments
            lit = int(rng.integers(0, 4))
            ml = int(rng.integers(4, 70))
            off = int(rng.integers(1, 4000))
        x += lit
        m = x
        for v in range(m, min(n, m + ml)):
            if off <= v:                       # source inside the tile
                d = v - m
                # a repeating match points into its first period (one hop, whatever its length)
                P[v] = (m - off) + (d % off) if off < ml and m - off >= 0 else v - off
            # else: the source lies before the tile: a root (its byte is final already)
        x += ml
    assert (P <= np.arange(n)).all()
    return P


def roots_of(P):
    R = P.copy()
    while True:
        R2 = R[R]
        if (R2 == R).all():
            return R
        R = R2


def resolve(P0, early_root):
    """returns (final P, rounds, pair-hops)"""
    P = P0.copy()
    n = len(P)
    assert n % 2 == 0
    x = np.arange(0, n, 2)
    pp = np.stack([P[x], P[x + 1]], axis=1)                        # the registers of the owners
    act = ~((pp[:, 0] == x) & (pp[:, 1] == x + 1))
    rounds = hops = 0
    while True:
        rounds += 1
        snap = P.copy()                                            # what the round reads
        hop = lambda w: np.stack([snap[w[:, 0]], snap[w[:, 1]]], axis=1)  # noqa: E731
        changed = np.zeros(len(x), dtype=bool)
        idx = np.nonzero(act)[0]
        if early_root:
            h1 = hop(pp[idx])
            hops += len(idx)
            same1 = (h1 == pp[idx]).all(axis=1)
            act[idx[same1]] = False
            go = idx[~same1]
            h2 = hop(h1[~same1])
            hops += len(go)
            pp[go] = h2
            changed[go] = True
            act[go[(h2 == h1[~same1]).all(axis=1)]] = False
        else:
            q = hop(hop(pp[idx]))
            hops += 2 * len(idx)
            mv = (q != pp[idx]).any(axis=1)
            pp[idx[mv]] = q[mv]
            changed[idx[mv]] = True
            act[idx[~mv]] = False
        P[x[changed]] = pp[changed, 0]                             # behind the barrier: owners store
        P[x[changed] + 1] = pp[changed, 1]
        if early_root:
            if not act.any():
                break
        elif not changed.any():
            break
        assert rounds < 64
    return P, rounds, hops


@pytest.mark.parametrize("kind", ["iid", "runs", "deep", "mixed"])
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_both_rules_reach_the_roots(kind, seed):
    rng = np.random.default_rng(seed)
    P0 = make_forest(rng, 8192, kind)
    want = roots_of(P0)
    a, rounds_a, hops_a = resolve(P0, early_root=False)
    b, rounds_b, hops_b = resolve(P0, early_root=True)
    assert (a == want).all()
    assert (b == want).all()
    assert rounds_b <= rounds_a
    assert hops_b <= hops_a


def test_identity_tile_takes_one_round_and_no_hops():
    P0 = np.arange(8192, dtype=np.int64)
    for rule in (False, True):
        P, rounds, hops = resolve(P0, early_root=rule)
        assert (P == P0).all() and rounds == 1 and hops == 0


def test_early_root_saves_about_a_round_on_flag_like_tiles():
    rng = np.random.default_rng(7)
    P0 = make_forest(rng, 8192, "iid")
    _, rounds_a, hops_a = resolve(P0, early_root=False)
    _, rounds_b, hops_b = resolve(P0, early_root=True)
    assert hops_b < 0.9 * hops_a, (rounds_a, hops_a, rounds_b, hops_b)
