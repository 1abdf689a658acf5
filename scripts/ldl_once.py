"""Build the KKT matrix of a workload (c2 | c4), factor twice and solve twice: the launch sequence of one refactor and
one LDL solve for ncu captures.  Usage: python scripts/ldl_once.py c2|c4"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import clarabel_rs_b200 as cb
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from helpers import workloads
which = sys.argv[1] if len(sys.argv) > 1 else "c2"
if which == "c2":
    pr = workloads.random_sparse_qp(n=100_000, m=200_000, nnz_per_row=5, seed=1, window=200)
else:
    pr = workloads.block_angular_qp(seed=3)
N, cp, rv, nz, ds = workloads.kkt_triu(pr["P"], pr["A"], np.random.default_rng(0).uniform(0.5, 2, pr["A"].shape[0]))
s = cb.CudaLDLSolver(N, cp, rv, nz, ds, ordering=cb.ORDER_ND)
for _ in range(2):
    assert s.refactor()
for _ in range(2):
    x = s.solve(np.ones(N))
print("ok")
