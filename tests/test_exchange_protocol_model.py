"""Exhaustive interleaving check of the fused counter-exchange protocol
(libflagstats_b200/csrc/flagstat_kernels.cuh, xchg_last_cta / xchg_collect): every rank's last CTA

    push   slot[peer][epoch & 1][me] = {my totals, tag(epoch)}     one store per peer and word
    take   slot[me][epoch & 1][r] for every r: re-read until the word carries tag(epoch),
           then use the data of THAT read                           -> the global counters

(every 8-byte word validates itself: there is no flag and no fence between data and flag), and a
rank starts epoch e+1 only after it finished e (serialised launches; with overlapped launches
the publishing part of e+1 still waits for the kernel of e, which is the same order).  The
model runs `world` such programs under EVERY sequentially-consistent interleaving (depth-first
over the reachable states).  Because check and data are one load, a reader can never use a
wrong epoch's data; what can go wrong is that the word it waits for is REPLACED by a later
epoch's before it looked -- then it waits forever (the kernel: FLAGSTAT_CUDA_ETIMEOUT).  So
the property checked is: no reachable state in which a rank is stuck.  Slots are
double-buffered by epoch parity, and that is sufficient because a rank cannot finish e+1 --
hence cannot push e+2 into the slot e used -- before every peer has pushed e+1, i.e. has
finished reading e.

The same checker must FIND the hang when the parity is taken away (single-buffered slots),
otherwise it proves nothing.  Memory-ordering (8-byte single-copy atomicity, the fence between
collecting and pushing) is outside this model; compute-sanitizer racecheck and the skewed-rank
GPU tests cover the real thing.  CPU only.
"""
import sys

import pytest


def _programs(world, epochs, buffers, modes=None, collect_first=True):
    """Per rank: a list of atomic steps (op, peer, epoch, buffer index).

    modes: one letter per epoch, "i" = immediate (push and take in the same launch,
    FLAGSTAT_cuda_device_allreduce) or "d" = deferred (push only; the NEXT launch -- or the
    trailing collect -- takes this epoch BEFORE it pushes its own,
    FLAGSTAT_cuda_device_allreduce_deferred / FLAGSTAT_cuda_xchg_collect).  collect_first=False
    models the wrong order (push own totals first, then collect the pending epoch)."""
    modes = modes or "i" * epochs
    assert len(modes) == epochs
    progs = []
    for me in range(world):
        steps = []
        pending = None

        def collect(e):
            for r in range(world):
                steps.append(("take", r, e, e % buffers))

        for e in range(1, epochs + 1):
            b = e % buffers
            steps.append(("work", None, e, b))  # counting the shard: no communication
            if pending is not None and collect_first:
                collect(pending)
                pending = None
            for r in range(world):
                steps.append(("push", r, e, b))
            if pending is not None:
                collect(pending)
                pending = None
            if modes[e - 1] == "i":
                collect(e)
            else:
                pending = e
        if pending is not None:
            collect(pending)
        progs.append(steps)
    return progs


def _explore(world, epochs, buffers, modes=None, collect_first=True):
    """Returns (states visited, a stuck state's (rank, peer, epoch waited for, epoch found) or None)."""
    progs = _programs(world, epochs, buffers, modes, collect_first)
    # memory: slot[owner][buffer][writer] = epoch whose words it holds
    zero = tuple(tuple(tuple(0 for _ in range(world)) for _ in range(buffers)) for _ in range(world))
    start = (tuple(0 for _ in range(world)), zero)
    seen = {start}
    stack = [start]
    stuck = None
    # work[me][pc] = shards rank `me` has finished counting when its program counter is pc
    work = []
    for prog in progs:
        acc, done_work = [0], 0
        for st in prog:
            done_work += st[0] == "work"
            acc.append(done_work)
        work.append(acc)
    _explore.max_lead = 0

    def put(mem, owner, b, writer, v):
        o = list(mem)
        bb = list(o[owner])
        w = list(bb[b])
        w[writer] = v
        bb[b] = tuple(w)
        o[owner] = tuple(bb)
        return tuple(o)

    while stack:
        pcs, slot = stack.pop()
        w = [work[me][pcs[me]] for me in range(world)]
        _explore.max_lead = max(_explore.max_lead, max(w) - min(w))
        moved = False
        blocked = []
        for me in range(world):
            pc = pcs[me]
            if pc == len(progs[me]):
                continue
            op, r, e, b = progs[me][pc]
            nslot = slot
            if op == "push":
                nslot = put(slot, r, b, me, e)
            elif op == "take" and slot[me][b][r] != e:
                blocked.append((me, r, e, slot[me][b][r]))
                continue
            moved = True
            npcs = pcs[:me] + (pc + 1,) + pcs[me + 1:]
            nxt = (npcs, nslot)
            if nxt not in seen:
                seen.add(nxt)
                stack.append(nxt)
        if not moved and blocked and stuck is None:
            # nobody can move: report the rank whose word was replaced, if there is one
            stuck = max(blocked, key=lambda t: t[3] - t[2])
    return len(seen), stuck


@pytest.mark.parametrize("world,epochs", [(2, 5), (3, 3)])
def test_parity_double_buffering_is_safe_under_every_interleaving(world, epochs):
    states, stuck = _explore(world, epochs, buffers=2)
    assert stuck is None, f"rank {stuck[0]} waits for epoch {stuck[2]} of rank {stuck[1]}, the slot holds {stuck[3]}"
    assert states > 150  # the search really branched


def test_the_checker_finds_the_overwrite_without_double_buffering():
    """Single-buffered slots: a fast rank's push of e+1 can replace its words of e before a slow
    rank has taken them; the slow rank then waits for a tag that never comes back (the kernel
    would turn that into FLAGSTAT_CUDA_ETIMEOUT).  The checker must report it."""
    states, stuck = _explore(2, 3, buffers=1)
    assert stuck is not None
    me, r, e, got = stuck
    assert got == e + 1  # the reader found its peer's NEXT epoch


def test_three_buffers_are_not_needed():
    # more buffers than parity gives are safe too, just unnecessary: same verdict
    s2, v2 = _explore(2, 4, buffers=2)
    s3, v3 = _explore(2, 4, buffers=3)
    assert v2 is None and v3 is None and s3 >= s2


@pytest.mark.parametrize("world,modes", [(2, "ddddd"), (3, "ddd"), (2, "didid"), (2, "ddiid"), (2, "iddii"), (3, "did")])
def test_deferred_collection_is_safe_under_every_interleaving(world, modes):
    """Deferred collection: a launch pushes only, its successor first collects the pending epoch and
    then pushes.  A rank may now be one whole epoch ahead of a peer; parity double buffering
    still suffices because a rank overwrites the slots of parity (e & 1) with epoch e + 2 only
    after it has collected e + 1, and a peer pushes e + 1 only after it has collected e.
    All-deferred, and every mixture with immediate calls."""
    states, stuck = _explore(world, len(modes), buffers=2, modes=modes)
    assert stuck is None, f"rank {stuck[0]} waits for epoch {stuck[2]} of rank {stuck[1]}, the slot holds {stuck[3]}"
    assert states > 150


def test_deferred_lets_a_rank_run_one_epoch_ahead():
    """What the deferred order buys.  With the wait in the same launch a rank can have counted at
    most ONE shard more than the slowest rank (it cannot leave epoch e before every peer has
    pushed e); deferred, it may have counted TWO more (its epoch e + 1 shard is done while a
    peer is still counting e: only the push of e + 1 waits for that peer's push of e).  That is
    the slack which keeps per-step jitter between GPUs off the critical path."""
    _explore(2, 4, buffers=2, modes="iiii")
    lead_i = _explore.max_lead
    _explore(2, 4, buffers=2, modes="dddd")
    lead_d = _explore.max_lead
    assert (lead_i, lead_d) == (1, 2)


def test_the_checker_finds_the_wrong_order_of_deferred_collection():
    """Pushing one's own totals BEFORE collecting the pending epoch breaks the argument: the
    push of e + 1 no longer implies that e has been read, so a fast peer can overwrite
    parity (e & 1) with e + 2 under a slow reader's nose.  The checker must see it."""
    _, stuck = _explore(2, 5, buffers=2, modes="ddddd", collect_first=False)
    assert stuck is not None


if __name__ == "__main__":
    print(_explore(int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])))
