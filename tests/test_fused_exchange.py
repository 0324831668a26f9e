"""Fused count + counter exchange (FLAGSTAT_cuda_device_allreduce): one kernel
launch per rank counts the shard and exchanges the 32 counters through
peer-mapped memory.  world = 1 runs on any GPU box; the multi-rank cases need
>= 2 GPUs (one process per GPU over CUDA IPC, and several ranks in ONE process
over peer access)."""
import ctypes as C
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


def test_world1_overwrite_and_accumulate(cuda_lib):
    import torch
    from libflagstats_b200 import sharded, synth
    from oracle import oracle as O

    x = sharded.FusedExchange()
    assert x.world == 1
    for n in (0, 1, 7, 1000, 16384 * 8 + 5, 3_000_003):
        d = synth.uniform_device(n + 1, 0, 11, 0x0FFF)[1:]
        want = O.flagstat_simd(O.synth_uniform(1, n, 11, 0x0FFF))
        out = torch.full((32,), 7, dtype=torch.int64, device="cuda")
        x.flagstat(d, out=out)  # overwrite: the 7s must be gone
        assert out.cpu().numpy().view(np.uint64).tolist() == want.tolist(), n
        x.flagstat(d, out=out, accumulate=True)
        assert out.cpu().numpy().view(np.uint64).tolist() == (2 * want).tolist(), n
    # raw pospopcnt through the same path
    d = synth.uniform_device(100_003, 0, 5, 0xFFFF)
    out = x.flagstat(d, pospopcnt=True)
    assert out.cpu().numpy().view(np.uint64).tolist() == O.pospopcnt(O.synth_uniform(0, 100_003, 5, 0xFFFF)).tolist()
    # samtools mode (exact n_pair_all in slots 0 / 16) through the same exchange
    a = O.synth_uniform(0, 100_003, 5, 0xFFFF)
    want = O.flagstat_simd(a)
    st = O.samtools_loop(a)
    want[0], want[16] = np.uint64(st[2, 0]), np.uint64(st[2, 1])
    out = x.flagstat(d, samtools=True)
    assert out.cpu().numpy().view(np.uint64).tolist() == want.tolist()
    assert sharded.flagstat_sharded(d, samtools=True).cpu().numpy().view(np.uint64).tolist() == want.tolist()
    # one rank: a deferred call has nobody to wait for and writes at once; collect is a no-op
    out = torch.full((32,), 7, dtype=torch.int64, device="cuda")
    x.flagstat(d, out=out, deferred=True)
    torch.cuda.synchronize()
    assert out.cpu().numpy().view(np.uint64).tolist() == O.flagstat_simd(a).tolist()
    x.collect()
    x.status()
    x.close()


def test_world1_overlapped_steps(cuda_lib):
    """Overlapped launches (programmatic dependent launch, FLAGSTAT_cuda_xchg_set_overlap):
    back-to-back calls on a non-default stream over different shards, accumulate and
    overwrite mixed, plain launches in between -- every result as if strictly serialised."""
    import torch
    from libflagstats_b200 import sharded, synth
    from oracle import oracle as O

    x = sharded.FusedExchange(overlap=True)
    sizes = [5_000_011, 16384 * 8 * 40, 1, 777_777, 30_000_001]
    shards = [synth.uniform_device(n, 1000 * i, 3 + i, 0x0FFF) for i, n in enumerate(sizes)]
    wants = [O.flagstat_simd(O.synth_uniform(1000 * i, n, 3 + i, 0x0FFF)) for i, n in enumerate(sizes)]
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    reps = 25
    with torch.cuda.stream(s):
        acc = torch.zeros(32, dtype=torch.int64, device="cuda")
        outs = []
        for rep in range(reps):
            for d in shards:
                x.flagstat(d, out=acc, accumulate=True, stream=s)
            o = torch.full((32,), -1, dtype=torch.int64, device="cuda")
            x.flagstat(shards[rep % len(shards)], out=o, accumulate=False, stream=s)
            outs.append(o)
            # plain accumulate launches between two collectives, serialised and overlapped
            cuda_lib.flagstat_device(shards[0], out=acc, stream=s, overlap=(rep % 2 == 0))
    s.synchronize()
    x.status()
    total = sum(wants) * np.uint64(reps) + wants[0] * np.uint64(reps)
    assert acc.cpu().numpy().view(np.uint64).tolist() == total.tolist()
    for rep, o in enumerate(outs):
        assert o.cpu().numpy().view(np.uint64).tolist() == wants[rep % len(shards)].tolist(), rep
    assert x.set_overlap(False) is True
    x.close()


def _worker(rank, world, port, n, reps, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from libflagstats_b200 import sharded, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    x = sharded.FusedExchange()
    lo, hi = sharded.shard_range(n, world, rank)
    local = synth.hiseqx_device(hi - lo, start=lo, seed=4, qcfail_ppm=5000, device=f"cuda:{rank}")
    outs = []
    out = torch.zeros(32, dtype=torch.int64, device=f"cuda:{rank}")
    # back-to-back collectives without host synchronisation: exercises the
    # epoch / double-buffer protocol (a fast rank runs ahead of a slow one)
    for i in range(reps):
        if i % 3 == rank % 3:
            torch.cuda._sleep(2_000_000)  # skew the ranks
        sharded.flagstat_sharded_fused(local, x, out=out)
        outs.append(out.clone())
    acc = torch.zeros(32, dtype=torch.int64, device=f"cuda:{rank}")
    for _ in range(3):
        sharded.flagstat_sharded_fused(local, x, out=acc, accumulate=True)
    # the samtools mode is one more collective in the same epoch sequence
    sam = x.flagstat(local, samtools=True)
    torch.cuda.synchronize()
    # overlapped steps on a non-default stream: same answers, skewed ranks
    x.set_overlap(True)
    side = torch.cuda.Stream()
    ov = []
    with torch.cuda.stream(side):
        acc2 = torch.zeros(32, dtype=torch.int64, device=f"cuda:{rank}")
        for i in range(reps):
            if i % 4 == rank % 4:
                torch.cuda._sleep(1_000_000)
            o = torch.full((32,), -1, dtype=torch.int64, device=f"cuda:{rank}")
            x.flagstat(local, out=o, stream=side)
            ov.append(o)
            x.flagstat(local, out=acc2, accumulate=True, stream=side)
    side.synchronize()
    same_ov = all(torch.equal(o, outs[0]) for o in ov) and torch.equal(acc2, outs[0] * reps)
    # deferred collection (push only; the next call or collect() writes the counters), still
    # overlapped and skewed, mixed with immediate calls and a collect in the middle
    dv = []
    with torch.cuda.stream(side):
        acc3 = torch.zeros(32, dtype=torch.int64, device=f"cuda:{rank}")
        for i in range(reps):
            if i % 3 == (rank + 1) % 3:
                torch.cuda._sleep(1_500_000)
            o = torch.full((32,), -1, dtype=torch.int64, device=f"cuda:{rank}")
            x.flagstat(local, out=o, stream=side, deferred=True)
            dv.append(o)
            x.flagstat(local, out=acc3, accumulate=True, stream=side, deferred=True)
            if i == reps // 2:
                x.collect(stream=side)
            if i % 5 == 4:
                o2 = torch.full((32,), -1, dtype=torch.int64, device=f"cuda:{rank}")
                x.flagstat(local, out=o2, stream=side)  # immediate: collects the pending one first
                dv.append(o2)
        x.collect(stream=side)
    side.synchronize()
    same_ov = same_ov and all(torch.equal(o, outs[0]) for o in dv) and torch.equal(acc3, outs[0] * reps)
    x.set_overlap(False)
    x.status()
    same = all(torch.equal(o, outs[0]) for o in outs) and same_ov
    q.put((rank, outs[0].cpu().numpy().view(np.uint64).tolist(), same,
           acc.cpu().numpy().view(np.uint64).tolist(), sam.cpu().numpy().view(np.uint64).tolist()))
    dist.barrier()
    x.close()
    dist.destroy_process_group()


def test_one_process_per_gpu_ipc_matches_oracle():
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    from oracle import oracle as O
    world = min(torch.cuda.device_count(), 8)
    n = 50_000_017
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, 12, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    a = O.synth_hiseqx(0, n, 4, 5000)
    want = O.numpy_flagstat(a)
    st = O.samtools_loop(a)
    want_sam = want.copy()
    want_sam[0], want_sam[16] = np.uint64(st[2, 0]), np.uint64(st[2, 1])
    for rank, first, same, acc, sam in got:
        assert first == want.tolist(), rank
        assert same, rank
        assert acc == (3 * want).tolist(), rank
        assert sam == want_sam.tolist(), rank


def test_ranks_in_one_process_peer_access(cuda_lib):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    from libflagstats_b200 import sharded, synth
    from oracle import oracle as O
    lib = cuda_lib.lib()
    world = min(torch.cuda.device_count(), 4)
    hs = (C.c_void_p * world)()
    for r in range(world):
        with torch.cuda.device(r):
            h = C.c_void_p()
            cuda_lib.check(lib.FLAGSTAT_cuda_xchg_create(C.byref(h), r, world, None), "create")
            hs[r] = h
    cuda_lib.check(lib.FLAGSTAT_cuda_xchg_connect_local(hs, world), "connect_local")
    n = 20_000_003
    shards, outs = [], []
    for r in range(world):
        lo, hi = sharded.shard_range(n, world, r)
        shards.append(synth.uniform_device(hi - lo, lo, 9, 0x0FFF, device=f"cuda:{r}"))
        outs.append(torch.zeros(32, dtype=torch.int64, device=f"cuda:{r}"))
    for rep in range(4):
        for r in range(world):
            with torch.cuda.device(r):
                st = torch.cuda.current_stream(r).cuda_stream
                cuda_lib.check(lib.FLAGSTAT_cuda_device_allreduce(
                    hs[r], shards[r].data_ptr(), shards[r].numel(), outs[r].data_ptr(), 0, st), "allreduce")
    want = O.flagstat_simd(O.synth_uniform(0, n, 9, 0x0FFF)).tolist()
    for r in range(world):
        torch.cuda.synchronize(r)
        assert outs[r].cpu().numpy().view(np.uint64).tolist() == want, r
        cuda_lib.check(lib.FLAGSTAT_cuda_xchg_status(hs[r]), "status")
    # deferred collection: three pushes per rank, each collected by the next call, the last by _collect
    acc = [torch.zeros(32, dtype=torch.int64, device=f"cuda:{r}") for r in range(world)]
    for rep in range(3):
        for r in range(world):
            with torch.cuda.device(r):
                st = torch.cuda.current_stream(r).cuda_stream
                cuda_lib.check(lib.FLAGSTAT_cuda_device_allreduce_deferred(
                    hs[r], shards[r].data_ptr(), shards[r].numel(), acc[r].data_ptr(), 1, st), "deferred")
    for r in range(world):
        with torch.cuda.device(r):
            cuda_lib.check(lib.FLAGSTAT_cuda_xchg_collect(hs[r], torch.cuda.current_stream(r).cuda_stream), "collect")
    for r in range(world):
        torch.cuda.synchronize(r)
        assert acc[r].cpu().numpy().view(np.uint64).tolist() == [3 * v for v in want], r
        cuda_lib.check(lib.FLAGSTAT_cuda_xchg_status(hs[r]), "status")
    for r in range(world):
        lib.FLAGSTAT_cuda_xchg_destroy(hs[r])


def test_missing_peer_times_out_instead_of_hanging(cuda_lib):
    """A rank whose peer never launches gives up after the timeout, leaves the
    output untouched and reports FLAGSTAT_CUDA_ETIMEOUT."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    from libflagstats_b200 import synth
    lib = cuda_lib.lib()
    hs = (C.c_void_p * 2)()
    for r in range(2):
        with torch.cuda.device(r):
            h = C.c_void_p()
            cuda_lib.check(lib.FLAGSTAT_cuda_xchg_create(C.byref(h), r, 2, None), "create")
            hs[r] = h
    cuda_lib.check(lib.FLAGSTAT_cuda_xchg_connect_local(hs, 2), "connect_local")
    cuda_lib.check(lib.FLAGSTAT_cuda_xchg_set_timeout_ms(hs[0], 200), "timeout")
    with torch.cuda.device(0):
        d = synth.uniform_device(100_000, 0, 1, 0x0FFF, device="cuda:0")
        out = torch.full((32,), -5, dtype=torch.int64, device="cuda:0")
        cuda_lib.check(lib.FLAGSTAT_cuda_device_allreduce(
            hs[0], d.data_ptr(), d.numel(), out.data_ptr(), 0, torch.cuda.current_stream(0).cuda_stream), "ar")
        torch.cuda.synchronize(0)
        assert lib.FLAGSTAT_cuda_xchg_status(hs[0]) == -5
        assert out.cpu().tolist() == [-5] * 32
        # the failure is sticky: the handle refuses further collectives (ESTATE) until recreated
        assert lib.FLAGSTAT_cuda_device_allreduce(
            hs[0], d.data_ptr(), d.numel(), out.data_ptr(), 0, torch.cuda.current_stream(0).cuda_stream) == -4
    # ... and the rank that gave up told its peer: rank 1 does not wait for a partner that is gone
    with torch.cuda.device(1):
        d1 = synth.uniform_device(100_000, 0, 1, 0x0FFF, device="cuda:1")
        out1 = torch.full((32,), -7, dtype=torch.int64, device="cuda:1")
        rc = lib.FLAGSTAT_cuda_device_allreduce(
            hs[1], d1.data_ptr(), d1.numel(), out1.data_ptr(), 0, torch.cuda.current_stream(1).cuda_stream)
        torch.cuda.synchronize(1)
        assert rc == 0 and lib.FLAGSTAT_cuda_xchg_status(hs[1]) == -5
        assert out1.cpu().tolist() == [-7] * 32
    for r in range(2):
        lib.FLAGSTAT_cuda_xchg_destroy(hs[r])
