"""Sharded LDL^T on N GPUs against the same factorisation on one GPU (round-2 measurement, not part of bench.py's
contract).  Launch:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1
                     --master-port 29511 scripts/shard_bench.py --workload c2|c4 [--reps 10]
Every rank builds the KKT matrix of the workload (h = 1), creates its rank of the sharded factorisation, and the job
times `refactor` (two phases + all-gather of the cut roots' update matrices) and `solve` (two forward phases + backward
+ all-gathers) with CUDA events, max over ranks; rank 0 also times the unsharded object and prints one JSON line."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    import clarabel_rs_b200 as cb
    from helpers import workloads
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--reps", type=int, default=10)
    a = ap.parse_args()
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    if a.workload == "c2":
        pr = workloads.random_sparse_qp(n=100_000, m=200_000, nnz_per_row=5, seed=1, window=200)
    else:
        pr = workloads.block_angular_qp(seed=3)
    N, cp, rv, nz, ds = workloads.kkt_triu(pr["P"], pr["A"], np.ones(pr["A"].shape[0]))
    t0 = time.perf_counter()
    s = cb.ShardedLDLRank(N, cp, rv, nz, ds, device=local, ordering=cb.ORDER_ND)
    t_create = time.perf_counter() - t0
    b = np.random.default_rng(0).standard_normal(N)

    def timed(fn, reps):
        fn()
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / reps], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ok = s.refactor()
    ms_ref = timed(s.refactor, a.reps)
    x = s.solve(b)
    ms_sol = timed(lambda: s.solve(b), a.reps)     # includes the H2D of b and the D2H of x of the convenience wrapper
    out = dict(workload=a.workload, n_gpus=world, N=int(N), ok=bool(ok), sharded_refactor_ms=ms_ref, sharded_solve_ms=ms_sol,
               create_s=t_create, exchange_doubles=[int(s._L.cldl_shard_count(s.solver._h, 0, r)) for r in range(world)],
               x_entries=[int(s._L.cldl_shard_count(s.solver._h, 2, r)) for r in range(world)])
    if rank == 0:
        one = cb.CudaLDLSolver(N, cp, rv, nz, ds, device=local, ordering=cb.ORDER_ND)
        one.refactor()
        out["single_refactor_ms"] = one.time_refactor_ms(a.reps)
        out["single_solve_ms"] = one.time_solve_ms(a.reps)
        x1 = one.solve(b)
        out["max_abs_diff_vs_single"] = float(np.max(np.abs(x - x1)))
        print(json.dumps(out))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
