"""N > 1 host logic on CPU: instance-index shards and the ensemble-statistics merge, world_size 2 over gloo.
(The step path has no collective; per-instance results do not depend on the shard layout -- the GPU tests
check that across batch sizes.)"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from slam_localization_b200 import fleet


def test_shard_ranges_partition_the_fleet():
    for total in (0, 1, 7, 64, 65536, 4194304 + 3):
        for world in (1, 2, 3, 4, 8):
            spans = [fleet.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        fleet.shard_range(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _FakeEngine:
    """Stands in for slam_localization_b200.engine on CPU: records what fleet.make_nccl_comm hands to NcclComm."""

    class NcclComm:
        def __init__(self, rank, world, exchange, device=None):
            ident = bytes(range(128)) if rank == 0 else None      # rank 0 "creates" the 128-byte unique id ...
            self.ident = exchange(ident)                          # ... and every rank must end up holding it
            self.rank, self.world = rank, world


def _worker(rank, world, port, total, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    comm = fleet.make_nccl_comm(_FakeEngine, rank, world)         # the id exchange of the C-level gather (slb_gather_stats)
    assert comm.ident == bytes(range(128)) and (comm.rank, comm.world) == (rank, world)
    x = np.random.default_rng(7).normal(size=(total, n))          # every rank can build the whole fleet ...
    lo, hi = fleet.shard_range(total, rank, world)
    local = torch.from_numpy(fleet.stats_from_vectors(x[lo:hi]))  # ... but only reduces its own shard
    merged = fleet.merge_stats(local)
    if rank == 0:
        q.put(merged.numpy().copy())
    dist.barrier()
    dist.destroy_process_group()


def test_ensemble_stats_merge_world_size_2_gloo():
    total, n, world = 1001, 12, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    merged = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    x = np.random.default_rng(7).normal(size=(total, n))
    ref = fleet.stats_from_vectors(x)
    np.testing.assert_allclose(merged, ref, rtol=1e-12, atol=1e-9)
    cnt, mean, cov = fleet.moments(merged, n)
    assert cnt == total
    np.testing.assert_allclose(mean, x.mean(axis=0), atol=1e-12)
    np.testing.assert_allclose(cov, np.cov(x.T, bias=True), atol=1e-10)


def test_merge_is_a_noop_without_process_group():
    s = torch.arange(5, dtype=torch.float64)
    assert torch.equal(fleet.merge_stats(s.clone()), s)
