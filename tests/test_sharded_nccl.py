"""One process per GPU over NCCL: contiguous range shards of one global column,
the CUDA kernel per shard, one all-reduce of the 32 counters.  Needs >= 2 GPUs
(skipped on a single-GPU box); world size = min(device_count, 8)."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import libflagstats_b200 as fs
    from libflagstats_b200 import sharded, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    lo, hi = sharded.shard_range(n, world, rank)
    # each rank regenerates ITS range of the global column in its own HBM
    local = synth.hiseqx_device(hi - lo, start=lo, seed=4, qcfail_ppm=5000, device=f"cuda:{rank}")
    out = sharded.flagstat_sharded(local)
    torch.cuda.synchronize()
    q.put((rank, out.cpu().numpy().view(np.uint64).tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_range_sharded_allreduce_matches_single_gpu_and_oracle():
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    import libflagstats_b200 as fs
    from libflagstats_b200 import synth
    from oracle import oracle as O
    world = min(torch.cuda.device_count(), 8)
    n = 50_000_017
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    single = fs.flagstat_u64(synth.hiseqx_device(n, 0, 4, 5000)).tolist()
    want = O.numpy_flagstat(O.synth_hiseqx(0, n, 4, 5000)).tolist()
    assert single == want
    for r in range(world):
        assert got[r] == want, r
    # several GPUs from ONE process through the C ABI (host pointer, host-side sum)
    a = O.synth_hiseqx(0, 20_000_003, 4, 5000)
    import ctypes as C
    f = np.zeros(32, np.uint64)
    fs.check(fs.lib().FLAGSTAT_cuda_multi_u64(a.ctypes.data, a.size,
                                              f.ctypes.data_as(C.POINTER(C.c_uint64)), world), "multi")
    assert f.tolist() == O.numpy_flagstat(a).tolist()
