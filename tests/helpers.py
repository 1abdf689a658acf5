import importlib.util
import os

import numpy as np

import clarabel_rs_b200 as cb

_wl = importlib.util.spec_from_file_location(
    "workloads", os.path.join(os.path.dirname(cb.pkg.__file__), "workloads.py"))
workloads = importlib.util.module_from_spec(_wl)
_wl.loader.exec_module(workloads)


def kkt_symv(N, cp, rv, nz, x):
    """y = K x for a triu-stored symmetric K (mirrors csc/matrix_math.rs:178-208)."""
    cols = np.repeat(np.arange(N), np.diff(cp))
    y = np.zeros(N)
    np.add.at(y, rv, nz * x[cols])
    off = rv != cols
    np.add.at(y, cols[off], nz[off] * x[rv[off]])
    return y


def small_kkt(n, m, seed, window=None, k=3):
    pr = workloads.random_sparse_qp(n=n, m=m, nnz_per_row=k, seed=seed, window=window, p_offdiag=n)
    rng = np.random.default_rng(seed + 100)
    h = rng.uniform(0.5, 2.0, size=m)
    return workloads.kkt_triu(pr["P"], pr["A"], h)
