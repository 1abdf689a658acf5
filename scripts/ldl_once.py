"""Build a C2-like KKT, factor a few times (for ncu captures of the LDL kernels)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import clarabel_rs_b200 as cb
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from helpers import workloads
n, m = int(sys.argv[1]), int(sys.argv[2])
window = None if sys.argv[3] == "None" else int(sys.argv[3])
pr = workloads.random_sparse_qp(n=n, m=m, nnz_per_row=5, seed=1, window=window)
N, cp, rv, nz, ds = workloads.kkt_triu(pr["P"], pr["A"], np.random.default_rng(0).uniform(0.5, 2, m))
s = cb.CudaLDLSolver(N, cp, rv, nz, ds, ordering=cb.ORDER_ND)
for _ in range(3):
    assert s.refactor()
x = s.solve(np.ones(N))
print("ok", s.time_refactor_ms(3))
print("solve_ms", s.time_solve_ms(20))
