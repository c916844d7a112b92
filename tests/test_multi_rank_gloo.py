"""CPU, world_size 2, gloo: the host-side logic of the multi-GPU path
(trajectory-range shards, max-over-ranks time, summed units) - the kernels are
not involved (there is no data-path collective)."""

import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hy_b200.shard import shard_bounds, reduce_throughput, device_for_iteration


def test_shard_bounds_cover_exactly():
    for total in (0, 1, 7, 125000, 1000003):
        for world in (1, 2, 3, 8):
            seen = 0
            prev = 0
            for r in range(world):
                lo, hi = shard_bounds(total, r, world)
                assert lo == prev and hi >= lo
                prev = hi
                seen += hi - lo
            assert seen == total and prev == total
    assert shard_bounds(1000000, 3, 8) == (375000, 500000)
    assert [device_for_iteration(i, 8) for i in range(10)] == [0, 1, 2, 3, 4, 5, 6, 7, 0, 1]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from hy_b200 import workloads as W

    lo, hi = shard_bounds(10, rank, world)
    # every rank generates ITS shard of the ensemble: seeds differ per rank
    ic = W.oss_ensemble(hi - lo, seed=20251019 + 7919 * rank)
    local_seconds = 1.0 + rank          # rank 1 is the slow one
    local_units = 100.0 * (hi - lo)
    val, tmax, units = reduce_throughput(local_seconds, local_units, dist)
    chk = torch.tensor([float(ic.sum())], dtype=torch.float64)
    gathered = [torch.zeros(1, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(gathered, chk)
    if rank == 0:
        out.put((val, tmax, units, [float(g.item()) for g in gathered]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_reduction_gloo():
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    val, tmax, units, sums = out.get()
    assert tmax == 2.0 and units == 1000.0 and val == 500.0
    assert sums[0] != sums[1]  # different shards, not replicas
