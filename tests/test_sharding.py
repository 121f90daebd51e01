"""Host-side sharding: pure-Python properties + a world_size-2 gloo run (the N>1 path on CPU)."""
import os
import random
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fbkst_b200 import sharding


def lengths_cfg3(n=500, seed=0):
    rng = random.Random(seed)
    return [rng.randint(200, 3000) for _ in range(n)]


def test_bucket_respects_budget_and_covers_everything():
    lens = lengths_cfg3()
    batches = sharding.bucket_by_length(lens, max_frames=96000)
    seen = sorted(i for b in batches for i in b)
    assert seen == list(range(len(lens)))
    for b in batches:
        assert len(b) * max(lens[i] for i in b) <= 96000
        assert [lens[i] for i in b] == sorted((lens[i] for i in b), reverse=True)  # collater order
    with pytest.raises(ValueError):
        sharding.bucket_by_length([10, 200], max_frames=100)


def test_steps_are_balanced():
    lens = lengths_cfg3(2000, seed=1)
    steps = sharding.shard_steps(sharding.bucket_by_length(lens, 96000), world=8)
    assert all(len(st) == 8 for st in steps)
    assert sharding.step_imbalance(lens, steps[:-1]) < 1.15  # every full step within 15 % of its mean


def _worker(rank, world, port, lens, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = sharding.batches_for_rank(lens, 96000, rank, world)
    frames = torch.tensor([float(sum(lens[i] for b in mine for i in b))])
    count = torch.tensor([float(sum(len(b) for b in mine))])
    steps = torch.tensor([float(len(mine))])
    smin = steps.clone()
    dist.all_reduce(frames)          # only bookkeeping crosses ranks: no data-path collective
    dist.all_reduce(count)
    dist.all_reduce(steps, op=dist.ReduceOp.MAX)
    dist.all_reduce(smin, op=dist.ReduceOp.MIN)
    ids = [i for b in mine for i in b]
    gathered = [None] * world
    dist.all_gather_object(gathered, ids)
    if rank == 0:
        q.put((frames.item(), count.item(), steps.item(), smin.item(), gathered))
    dist.destroy_process_group()


def test_two_rank_gloo_partition():
    lens = lengths_cfg3(300, seed=2)
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, lens, q)) for r in range(2)]
    for p in procs:
        p.start()
    frames, count, smax, smin, gathered = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert frames == float(sum(lens)) and count == float(len(lens))
    assert smax == smin                                   # same number of steps on every rank
    assert sorted(gathered[0] + gathered[1]) == list(range(len(lens)))
    assert not set(gathered[0]) & set(gathered[1])        # disjoint shards
