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
