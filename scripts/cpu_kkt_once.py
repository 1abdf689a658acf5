"""CPU side of the north-star comparison on one configuration (test / measurement infrastructure: it runs oracle/).
The reference's single-thread qdldl (oracle port of qdldl.rs:469-669, :708-768) on the KKT matrix of the workload at
unit scaling: constructor (symbolic + first numeric factorisation), `refactor` and `solve`, each timed with
perf_counter, median of --reps, one thread (pin it: `taskset -c 2 python scripts/cpu_kkt_once.py ...`).
  --order amd   the ordering the reference itself would use (AMD, dense scale 1.5 -- ours stands in for the crate)
  --order nd    the nested-dissection ordering the GPU path uses (cheaper for the CPU too on C4: the conservative baseline)
"""
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
    import clarabel_rs_b200 as cb
    import oracle
    from helpers import workloads
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c4")
    ap.add_argument("--order", default="nd")
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    if a.workload == "c2":
        pr = workloads.random_sparse_qp(n=100_000, m=200_000, nnz_per_row=5, seed=1, window=200)
    else:
        pr = workloads.block_angular_qp(seed=3)
    N, cp, rv, nz, ds = workloads.kkt_triu(pr["P"], pr["A"], np.ones(pr["A"].shape[0]))
    t0 = time.perf_counter()
    perm = cb.order(N, cp, rv, cb.ORDER_AMD if a.order == "amd" else cb.ORDER_ND, 1.5)
    t_order = time.perf_counter() - t0
    t0 = time.perf_counter()
    f = oracle.QDLDL((N, N), cp, rv, nz, perm, dsigns=ds, regularize_eps=1e-13, regularize_delta=2e-7)
    t_new = time.perf_counter() - t0
    tr, ts = [], []
    b = np.random.default_rng(0).standard_normal(N)
    for _ in range(a.reps):
        t0 = time.perf_counter(); f.refactor(); tr.append(time.perf_counter() - t0)
    x = None
    for _ in range(max(a.reps, 5)):
        xb = b.copy()
        t0 = time.perf_counter(); f._L.oq_solve(f._h, oracle.P(xb)); ts.append(time.perf_counter() - t0)
    cpu = ""
    try:
        cpu = [l.split(":", 1)[1].strip() for l in open("/proc/cpuinfo") if l.startswith("model name")][0]
    except Exception:
        pass
    print(json.dumps(dict(workload=pr["name"], order=a.order, kkt_dim=int(N), nnzK=int(cp[-1]), nnzL=f.nnzL,
                          order_s=t_order, new_s=t_new, refactor_s=float(np.median(tr)), refactor_all=tr,
                          solve_s=float(np.median(ts)), regularize_count=f.regularize_count, threads=1, cpu=cpu,
                          affinity=sorted(os.sched_getaffinity(0)), host_cores=os.cpu_count())))


if __name__ == "__main__":
    main()
