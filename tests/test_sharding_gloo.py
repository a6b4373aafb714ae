"""Host-side logic of the N>1 path on CPU: contiguous environment sharding and the optional
gather epilogue, exercised with world_size 2 and 3 over the gloo backend (uneven shards too)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sapien_b200 import sharding


@pytest.mark.parametrize("n,world", [(1024, 8), (1024, 1), (64, 4), (10, 4), (3, 8), (0, 2), (7, 7)])
def test_env_range_is_a_contiguous_balanced_partition(n, world):
    ranges = [sharding.env_range(n, r, world) for r in range(world)]
    assert ranges[0][0] == 0 and ranges[-1][1] == n
    for (a0, a1), (b0, b1) in zip(ranges, ranges[1:]):
        assert a1 == b0 and a0 <= a1
    counts = sharding.env_counts(n, world)
    assert sum(counts) == n and max(counts) - min(counts) <= 1
    for env in range(n):
        r = sharding.owner_of(env, n, world)
        assert ranges[r][0] <= env < ranges[r][1]


def test_env_range_rejects_bad_ranks():
    with pytest.raises(ValueError):
        sharding.env_range(8, 2, 2)
    with pytest.raises(ValueError):
        sharding.env_range(8, 0, 0)


class FakeEngine:
    """Stands in for the CUDA engine in the gloo test (no GPU here): depth[n] = mean(left[n])."""

    def __init__(self, *args, device=-1, batch=1):
        self.batch = batch
        self.out = None

    def synchronize(self):
        pass

    def compute(self, left, right, sync=True):
        assert left.shape[0] == self.batch
        self.out = (left.float().mean(dim=(1, 2))[:, None, None] + right.float()).contiguous()

    def get_cuda(self):
        return self.out


def _worker(rank, world, port, n_envs, results, pipelines=1):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        gen = torch.Generator().manual_seed(1234)
        left = torch.randint(0, 256, (n_envs, 6, 5), generator=gen, dtype=torch.uint8)
        right = torch.randint(0, 256, (n_envs, 6, 5), generator=gen, dtype=torch.uint8)
        sh = sharding.ShardedStereoDepth((), n_envs, rank, world, engine_factory=FakeEngine, pipelines=pipelines)
        assert sum(b - a for a, b in sh.blocks) == sh.local and len(sh.engines) == min(pipelines, sh.local)
        local = sh.compute(left[sh.start:sh.stop], right[sh.start:sh.stop])  # None on a rank without environments
        assert (local is None) == (sh.local == 0)
        like = ((6, 5), torch.float32, "cpu")
        full = sh.gather_depth(local, like=like)  # all-gather
        root = sh.gather_depth(local, dst=0, like=like)
        expect = left.float().mean(dim=(1, 2))[:, None, None] + right.float()
        ok = torch.equal(full, expect) and ((rank != 0 and root is None) or (rank == 0 and torch.equal(root, expect)))
        # timing reduction used by bench.py: max over ranks
        t = torch.tensor([float(rank + 1)], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ok = ok and t.item() == world
        results[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world,n_envs,pipelines", [(2, 8, 1), (2, 5, 2), (3, 7, 1), (2, 2, 1), (3, 2, 3), (2, 1, 1)])
def test_sharded_compute_and_gather_over_gloo(world, n_envs, pipelines):
    """Even and uneven shards, sub-block pipelines, ranks that own exactly one environment (2, 2) and ranks that own
    none (3 ranks / 2 envs, 2 ranks / 1 env): those join the gather with an empty tensor."""
    port = _free_port()
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(world, port, n_envs, results, pipelines), nprocs=world, join=True)
    assert dict(results) == {r: True for r in range(world)}


def test_single_process_gather_is_identity():
    x = torch.arange(12.0).reshape(4, 3)
    assert sharding.gather_envs(x, 4) is x
