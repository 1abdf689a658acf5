"""CPU, world_size 2, gloo: the N>1 path of bench.py -- one independent problem per
rank (seed + rank), no data-path collective, whole-job value = sum of units / max time."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import bench
    import clarabel_rs_b200 as cb
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pr, desc = bench.load_workload("c2small", rank)
    # host-side work a rank does before touching its GPU: KKT pattern + symbolic analysis
    from helpers import workloads
    N, cp, rv, nz, ds = workloads.kkt_triu(pr["P"], pr["A"], np.ones(pr["A"].shape[0]))
    S = cb.SymbolicAnalysis(N, cp, rv, ordering=cb.ORDER_ND)
    steps, secs = 10, 0.5 * (rank + 1)          # rank 1 is the slow one
    value, tmax, ktot = bench.aggregate_over_ranks(dist, world, steps, secs)
    seeds = [None] * world
    dist.all_gather_object(seeds, desc)
    if rank == 0:
        out.put((value, tmax, ktot, seeds, S.nlevels))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_aggregation_and_sharding():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    value, tmax, ktot, seeds, nlev = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ktot == 20 and tmax == 1.0 and abs(value - 20.0) < 1e-12   # sum of steps / max time
    assert seeds[0] != seeds[1] and "seed=1" in seeds[0] and "seed=2" in seeds[1]   # independent problems
    assert nlev > 0
