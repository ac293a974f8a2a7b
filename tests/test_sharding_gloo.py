"""CPU, world_size 2, gloo: the N>1 host logic — chain sharding and the final gather of isotherm averages."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "mc-porousmaterials_b200"))


def fake_stats(chain):
    class S:
        pass
    s = S()
    s.samples = 1000 + chain
    s.n = 10 + chain
    s.sum_n = float((10 + chain) * s.samples)
    s.sum_n2 = float((10 + chain) ** 2 * s.samples)
    s.sum_e = -5.0 * chain * s.samples
    s.sum_e2 = (5.0 * chain) ** 2 * s.samples
    return s


def _worker(rank, world, n_chains, port, q):
    from mcpm_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    own = sharding.shard_chains(n_chains, rank, world)
    local = sharding.pack_accumulators([fake_stats(c) for c in own])
    full = sharding.gather_accumulators(dist, local, n_chains, rank, world)
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)           # the bench's max-over-ranks of the device time
    dist.barrier()
    if rank == 0:
        q.put((full, float(t.item())))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_chains", [8, 7])
def test_shard_and_gather_world2(n_chains):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 1000) + n_chains
    procs = [ctx.Process(target=_worker, args=(r, 2, n_chains, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    full, tmax = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert tmax == 2.0
    from mcpm_b200 import sharding
    want = sharding.pack_accumulators([fake_stats(c) for c in range(n_chains)])
    assert np.array_equal(full, want)
    # every chain is owned exactly once
    owners = sorted(sharding.shard_chains(n_chains, 0, 2) + sharding.shard_chains(n_chains, 1, 2))
    assert owners == list(range(n_chains))


def test_isotherm_reduction():
    from mcpm_b200 import sharding
    n_points, reps = 4, 3
    acc = sharding.pack_accumulators([fake_stats(c) for c in range(n_points * reps)])
    mean, sem, energy = sharding.isotherm_from_accumulators(acc, n_points)
    for p in range(n_points):
        vals = [10 + (r * n_points + p) for r in range(reps)]
        assert mean[p] == pytest.approx(np.mean(vals))
        assert sem[p] == pytest.approx(np.std(vals, ddof=1) / np.sqrt(reps))
