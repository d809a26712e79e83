"""CPU (gloo, world_size 2): stream sharding + result gather used on the multi-GPU path."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_shard_ranges_cover_streams():
    from bpv.dist import shard_range
    for S in (1, 2, 7, 256, 65536):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(S, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == S
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, S, K, ragged, q):
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(here, 'bp-from-video_b200'))
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from bpv import dist as bd
    r, _, w = bd.init_from_env('gloo')
    assert (r, w) == (rank, world)
    full = torch.arange(S * K, dtype=torch.float64).reshape(S, K)      # the "single-GPU" result
    lo, hi = bd.shard_range(S, rank, world)
    counts = [bd.shard_range(S, i, world)[1] - bd.shard_range(S, i, world)[0] for i in range(world)]
    got = bd.gather_records(full[lo:hi].clone(), counts if ragged else None)
    q.put((rank, bool(torch.equal(got, full))))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('S,ragged', [(8, False), (7, True)])
def test_gather_equals_single_rank_result(S, ragged):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 200) + S
    procs = [ctx.Process(target=_worker, args=(r, 2, port, S, 6, ragged, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert all(ok for _, ok in res), res


def _worker_async(rank, world, port, J, q):
    """RecordGather over gloo: the gather launched at step k is handed over at step k + 1 (int32 24-byte records)."""
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(here, 'bp-from-video_b200'))
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from bpv import dist as bd
    bd.init_from_env('gloo')
    rg = bd.RecordGather()
    ok = rg.collect() is None and rg.flush() is None
    expect = lambda k: torch.cat([torch.full((J, 6), 1000 * r + k, dtype=torch.int32) for r in range(world)])
    for k in range(5):
        rg.launch(torch.full((J, 6), 1000 * rank + k, dtype=torch.int32))
        prev = rg.collect()
        ok &= (prev is None) if k == 0 else bool(torch.equal(prev, expect(k - 1)))
    ok &= bool(torch.equal(rg.flush(), expect(4)))
    q.put((rank, ok))
    dist.barrier()
    dist.destroy_process_group()


def test_record_gather_hands_results_over_one_step_later():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29850 + (os.getpid() % 100)
    procs = [ctx.Process(target=_worker_async, args=(r, 2, port, 5, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert all(ok for _, ok in res), res


def test_record_gather_passes_through_without_a_process_group():
    from bpv.dist import RecordGather, _parse_cpulist
    rg = RecordGather()
    a, b = torch.arange(12).reshape(2, 6), torch.arange(12, 24).reshape(2, 6)
    rg.launch(a)
    assert rg.collect() is None
    rg.launch(b)
    assert torch.equal(rg.collect(), a) and torch.equal(rg.flush(), b)
    assert _parse_cpulist('0-3,8,10-11\n') == {0, 1, 2, 3, 8, 10, 11} and _parse_cpulist('') == set()
