"""Device timing probe for the LDL level (refactor / solve) on a C2-like KKT."""
import json, sys, time
import numpy as np
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import clarabel_rs_b200 as cb
sys.path.insert(0, "tests")
from helpers import workloads, kkt_symv

n, m, window = int(sys.argv[1]), int(sys.argv[2]), (None if sys.argv[3] == "None" else int(sys.argv[3]))
pr = workloads.random_sparse_qp(n=n, m=m, nnz_per_row=5, seed=1, window=window)
N, cp, rv, nz, ds = workloads.kkt_triu(pr["P"], pr["A"], np.random.default_rng(0).uniform(0.5, 2, m))
for name, o, mp in [("amd", cb.ORDER_AMD, 0), ("nd", cb.ORDER_ND, 0), ("nd32", cb.ORDER_ND, 32)]:
    t = time.time()
    s = cb.CudaLDLSolver(N, cp, rv, nz, ds, ordering=o, max_panel=mp)
    tc = time.time() - t
    ok = s.refactor()
    i = s.linear_solver_info()
    tr = s.time_refactor_ms(5)
    ts = s.time_solve_ms(10)
    b = np.random.default_rng(1).standard_normal(N)
    x = s.solve(b)
    res = np.max(np.abs(kkt_symv(N, cp, rv, nz, x) - b)) / np.max(np.abs(b))
    print(json.dumps(dict(order=name, create_s=round(tc, 2), ok=ok, refactor_ms=round(tr, 3), solve_ms=round(ts, 3),
                          nnzL=i.nnzL, stored=i.nnzL_stored, flops=i.flops, levels=i.n_levels, nsup=i.n_supernodes,
                          res=res, gflops=round(i.flops / tr / 1e6, 1), solve_gbs=round(16 * i.nnzL_stored / ts / 1e6, 1))), flush=True)
    s.close()
