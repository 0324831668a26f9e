"""World-size-2 gloo test of the multi-GPU host logic: contiguous range shards
+ the single all-reduce of the 32 counters.  The per-shard counters come from
the oracle here (no GPU in this container); on the B200 box the same
shard_range / allreduce_counters code runs over NCCL with the CUDA kernel."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, q):
    sys.path.insert(0, ROOT)
    from oracle import oracle as O
    from libflagstats_b200.sharded import allreduce_counters, shard_range
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(n, world, rank)
    local = O.flagstat_simd(O.synth_uniform(lo, hi - lo, 17, 0x0FFF))  # shard regenerated in place
    t = torch.from_numpy(local.view(np.int64).copy())
    allreduce_counters(t)
    q.put((rank, t.numpy().view(np.uint64).tolist()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [1_000_003, 64])
def test_two_rank_range_shard_allreduce(n):
    from oracle import oracle as O
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = O.flagstat_simd(O.synth_uniform(0, n, 17, 0x0FFF)).tolist()
    assert got[0] == want and got[1] == want


def test_allreduce_is_a_noop_without_a_group():
    from libflagstats_b200.sharded import allreduce_counters
    t = torch.arange(32, dtype=torch.int64)
    assert allreduce_counters(t.clone()).tolist() == t.tolist()
    with pytest.raises(ValueError):
        allreduce_counters(torch.zeros(32, dtype=torch.int32))
