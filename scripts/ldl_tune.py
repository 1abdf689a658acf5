"""Sweep of the symbolic knobs (nd_leaf, relax_subtree, max_panel) on a C2-like KKT: refactor / solve ms."""
import os, sys, subprocess, json
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
if len(sys.argv) > 1 and sys.argv[1] == "one":
    sys.path.insert(0, os.path.dirname(HERE))
    import clarabel_rs_b200 as cb
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), "tests"))
    from helpers import workloads
    nd_leaf, max_panel = int(sys.argv[2]), int(sys.argv[3])
    pr = workloads.random_sparse_qp(n=100000, m=200000, nnz_per_row=5, seed=1, window=200)
    N, cp, rv, nz, ds = workloads.kkt_triu(pr["P"], pr["A"], np.random.default_rng(0).uniform(0.5, 2, 200000))
    s = cb.CudaLDLSolver(N, cp, rv, nz, ds, ordering=cb.ORDER_ND, nd_leaf=nd_leaf, max_panel=max_panel)
    assert s.refactor()
    i = s.linear_solver_info()
    print(json.dumps(dict(nd_leaf=nd_leaf, max_panel=max_panel, relax=os.environ.get("CB_RELAX_SUBTREE", "32"),
                          refactor_ms=round(s.time_refactor_ms(5), 3), solve_ms=round(s.time_solve_ms(10), 3),
                          levels=i.n_levels, nsup=i.n_supernodes, nnzL_stored=i.nnzL_stored, gflop=round(i.flops / 1e9, 1))), flush=True)
else:
    for relax in ("16", "32", "64"):
        for nd_leaf in (100, 200, 400):
            env = dict(os.environ, CB_RELAX_SUBTREE=relax)
            subprocess.run([sys.executable, os.path.abspath(__file__), "one", str(nd_leaf), "0"], env=env, timeout=120)
    for mp in (32, 48):
        subprocess.run([sys.executable, os.path.abspath(__file__), "one", "200", str(mp)], timeout=120)
