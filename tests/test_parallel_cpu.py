"""Multi-process (gloo, world_size 2) test of the sharding + gather plumbing used on N > 1 GPUs."""
import os

import numpy as np
import pytest


def _worker(rank, world, port, tmp):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch
    import torch.distributed as dist
    import chd
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    B = 5
    ps = chd.synth.make_batch(B, 40, 2)
    full = chd.phys.PhysBatch(ps, host_only=True)
    shards = chd.parallel.shard_by_work(full.sizes[:, 0], world)
    slots = chd.parallel.pad_to(shards)
    mine = shards[rank]
    local = chd.phys.PhysBatch([ps[i] for i in mine], host_only=True)
    # stand-in for the sampled solution block: the initial iterate, padded to a fixed width
    width = 1500
    x = np.zeros((slots, 1, width))
    x0 = local.get_x()
    x[:len(mine), 0, :x0.shape[1]] = x0
    g = chd.parallel.gather_samples(torch.from_numpy(x), world).numpy()
    out = chd.parallel.unshard(g, shards, slots)
    if rank == 0:
        np.save(os.path.join(tmp, "gathered.npy"), out)
        np.save(os.path.join(tmp, "ref.npy"), full.get_x())
    dist.destroy_process_group()


def test_shard_and_gather_gloo(tmp_path):
    import torch.multiprocessing as mp
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    got = np.load(str(tmp_path / "gathered.npy"))
    ref = np.load(str(tmp_path / "ref.npy"))
    for i in range(ref.shape[0]):
        n = ref.shape[1]
        np.testing.assert_array_equal(got[i, 0, :n][ref[i] != 0], ref[i][ref[i] != 0])


def test_shard_by_work_balanced(chd):
    work = [10, 1, 9, 2, 8, 3, 7, 4]
    s = chd.parallel.shard_by_work(work, 2)
    assert sorted(s[0] + s[1]) == list(range(8))
    assert abs(sum(work[i] for i in s[0]) - sum(work[i] for i in s[1])) <= 2
    assert chd.parallel.shard_by_work([5, 5, 5], 4)[3] == []
