"""GPU parity tests of the device LDL^T through the C-ABI (cldl_*), against the
CPU oracle on the SAME permutation, and against the reference's own KATs."""
import numpy as np
import pytest

import clarabel_rs_b200 as cb
from helpers import kkt_symv, small_kkt, workloads
from oracle import QDLDL

pytestmark = pytest.mark.gpu

REL_TOL = 1e-9  # BASELINE.json north_star: 1e-9 relative on the KKT residual


def matrix_4x4():
    return 4, [0, 1, 3, 6, 8], [0, 0, 1, 0, 1, 2, 2, 3], [8., -3., 8., 2., -1., 8., -1., 1.]


@pytest.mark.parametrize("perm", [None, [0, 1, 2, 3], [3, 0, 1, 2], [3, 0, 2, 1]])
def test_qdldl_kat_4x4(perm):  # qdldl/test.rs:194-230
    n, Ap, Ai, Ax = matrix_4x4()
    s = cb.CudaLDLSolver(n, Ap, Ai, Ax, np.ones(4, np.int8), perm=perm,
                         regularize_eps=1e-12, regularize_delta=1e-7)
    assert s.refactor()
    x = s.solve([20.0, -22.0, 32.0, -7.0])
    assert np.max(np.abs(x - [1., -2., 3., -4.])) <= 1e-8


def test_solve_before_refactor_is_an_error():  # qdldl/test.rs:232-247 (reference panics)
    n, Ap, Ai, Ax = matrix_4x4()
    s = cb.CudaLDLSolver(n, Ap, Ai, Ax, np.ones(4, np.int8))
    with pytest.raises(cb.BackendError):
        s.solve([1., 2., 3., 4.])


def test_zero_pivot_without_regularisation():  # qdldl/test.rs:266-283
    n, Ap, Ai, Ax = matrix_4x4()
    A0 = list(Ax); A0[0] = 0.0
    s = cb.CudaLDLSolver(n, Ap, Ai, A0, np.ones(4, np.int8), perm=[0, 1, 2, 3], regularize_enable=False)
    with pytest.raises(cb.BackendError) as e:
        s.refactor()
    assert "ZeroPivot" in str(e.value)


def test_structural_errors():  # qdldl/test.rs:285-318
    with pytest.raises(cb.BackendError) as e:
        cb.CudaLDLSolver(3, [0, 3, 6, 9], [0, 1, 2] * 3, [1., 2., 1., 3., 3., 4., 5., 6., 7.], [1, 1, 1])
    assert "NotUpperTriangular" in str(e.value)
    with pytest.raises(cb.BackendError) as e:
        cb.CudaLDLSolver(3, [0, 1, 1, 3], [0, 0, 2], [1., 5., 7.], [1, 1, 1])
    assert "EmptyColumn" in str(e.value)
    with pytest.raises(cb.BackendError) as e:
        cb.CudaLDLSolver(3, [0, 1, 2, 3], [0, 1, 2], [1., 5., 7.], [1, 1, 1], perm=[0, 0, 1])
    assert "InvalidPermutation" in str(e.value)


@pytest.mark.parametrize("perm", [None, list(range(6)), [5, 4, 3, 2, 1, 0]])
def test_trait_level_golden(perm):  # ldlsolvers/faer_ldl.rs:352-409
    KKT = (6, [0, 1, 2, 4, 6, 8, 10], [0, 1, 0, 2, 1, 3, 0, 4, 1, 5],
           [1.0, 2.0, 1.0, -1.0, 1.0, -2.0, -1.0, -3.0, -1.0, -4.0])
    s = cb.CudaLDLSolver(*KKT, [1, 1, -1, -1, -1, -1], perm=perm)
    assert s.refactor()
    b = [1.0, 2.0, 3.0, 4., 5., 6.]
    x = s.solve(b)
    assert np.max(np.abs(x - [1.0, 0.9090909090909091, -2.0, -1.5454545454545454, -2.0,
                              -1.7272727272727275])) < 1e-10
    s.update_values([9], [-10.0])
    assert s.refactor()
    x = s.solve(b)
    assert np.max(np.abs(x - [1.0, 1.3076923076923077, -2.0, -1.346153846153846, -2.0,
                              -0.7307692307692306])) < 1e-10
    # offset/scale, cross-checked against the oracle doing the same thing
    s.offset_values([1, 2], 3., [1, -1])
    s.scale_values([1, 2], 2.)
    f = QDLDL((6, 6), KKT[1], KKT[2], KKT[3], s.perm(), dsigns=[1, 1, -1, -1, -1, -1],
              regularize_eps=1e-13, regularize_delta=2e-7)
    f.update_values([9], [-10.0]); f.offset_values([1, 2], 3., [1, -1]); f.scale_values([1, 2], 2.)
    f.refactor()
    assert s.refactor()
    assert np.max(np.abs(s.solve(b) - f.solve(b))) < 1e-10
    info = s.linear_solver_info()
    assert info.name == "cudaldl" and info.direct and info.nnzA == 10


CASES = [
    # n, m, window, ordering, max_panel
    (30, 50, None, cb.ORDER_AMD, 0),
    (200, 350, 20, cb.ORDER_ND, 0),
    (400, 300, None, cb.ORDER_BEST, 0),
    (400, 300, None, cb.ORDER_AMD, 8),       # forces chains of split panels
    (3000, 6000, 60, cb.ORDER_BEST, 0),
    (3000, 6000, 60, cb.ORDER_AMD, 0),
    (1500, 2500, None, cb.ORDER_AMD, 0),     # expander: big dense fronts (global-memory panel path)
    (1500, 2500, None, cb.ORDER_ND, 128),
]


@pytest.mark.parametrize("n,m,window,ordering,max_panel", CASES)
def test_parity_with_oracle(n, m, window, ordering, max_panel):
    N, cp, rv, nz, ds = small_kkt(n, m, seed=n + m, window=window, k=4)
    s = cb.CudaLDLSolver(N, cp, rv, nz, ds, ordering=ordering, max_panel=max_panel, nd_leaf=64)
    assert s.refactor()
    perm = s.perm()
    f = QDLDL((N, N), cp, rv, nz, perm, dsigns=ds, regularize_eps=1e-13, regularize_delta=2e-7)
    info = s.linear_solver_info()
    assert info.nnzL == f.nnzL
    assert info.regularize_count == f.regularize_count
    assert info.positive_inertia == f.positive_inertia == n
    rng = np.random.default_rng(7)
    for _ in range(2):
        b = rng.standard_normal(N)
        x, xo = s.solve(b), f.solve(b)
        scale = max(1.0, np.max(np.abs(xo)))
        assert np.max(np.abs(x - xo)) <= REL_TOL * scale
        # KKT residual parity (the north-star criterion)
        r, ro = kkt_symv(N, cp, rv, nz, x) - b, kkt_symv(N, cp, rv, nz, xo) - b
        nb = np.max(np.abs(b))
        assert np.max(np.abs(r)) / nb <= np.max(np.abs(ro)) / nb + REL_TOL
    # value update + refactor path, twice (checks arena reuse across refactors)
    for it in range(2):
        # only touch entries of the A' block (rows < n, columns >= n): K stays quasidefinite
        cols = np.repeat(np.arange(N), np.diff(cp))
        cand = np.nonzero((cols >= n) & (rv < n))[0]
        idx = rng.choice(cand, size=min(cand.size, 500), replace=False)
        vals = rng.standard_normal(idx.size)
        s.update_values(idx, vals); f.update_values(idx, vals)
        nz = nz.copy(); nz[idx] = vals
        assert s.refactor(); f.refactor()
        b = rng.standard_normal(N)
        x, xo = s.solve(b), f.solve(b)
        assert np.max(np.abs(x - xo)) <= 1e-8 * max(1.0, np.max(np.abs(xo)))
        assert s.linear_solver_info().regularize_count == f.regularize_count


def test_dynamic_regularisation_counts():
    # diagonal entries that violate the expected sign get delta*sign (qdldl.rs:645-651)
    N, cp, rv, nz, ds = small_kkt(100, 150, seed=5, window=10)
    nz = nz.copy()
    dgl = cp[1:] - 1
    nz[dgl[100:130]] = 0.0          # zero Hs entries (zero cone rows): pivot becomes -(a' P^-1 a) < 0, fine
    nz[dgl[:5]] = -1.0              # wrong-signed P diagonal: forces regularisation
    s = cb.CudaLDLSolver(N, cp, rv, nz, ds, ordering=cb.ORDER_AMD)
    assert s.refactor()
    f = QDLDL((N, N), cp, rv, nz, s.perm(), dsigns=ds, regularize_eps=1e-13, regularize_delta=2e-7)
    assert s.linear_solver_info().regularize_count == f.regularize_count > 0
    b = np.random.default_rng(1).standard_normal(N)
    x, xo = s.solve(b), f.solve(b)
    assert np.max(np.abs(x - xo)) <= 1e-7 * max(1.0, np.max(np.abs(xo)))


def test_reproducible_bitwise():
    N, cp, rv, nz, ds = small_kkt(2000, 3000, seed=11, window=40)
    s = cb.CudaLDLSolver(N, cp, rv, nz, ds)
    b = np.random.default_rng(2).standard_normal(N)
    s.refactor(); x1 = s.solve(b)
    s.refactor(); x2 = s.solve(b)
    assert np.array_equal(x1, x2)


def test_full_size_roundtrip_property():
    """BASELINE config C2 size (N = 3e5): size-independent property K x = b."""
    pr = workloads.random_sparse_qp(n=100_000, m=200_000, nnz_per_row=5, seed=1, window=200)
    rng = np.random.default_rng(3)
    N, cp, rv, nz, ds = workloads.kkt_triu(pr["P"], pr["A"], rng.uniform(0.5, 2.0, size=200_000))
    s = cb.CudaLDLSolver(N, cp, rv, nz, ds)
    assert s.refactor()
    xt = rng.standard_normal(N)
    b = kkt_symv(N, cp, rv, nz, xt)
    x = s.solve(b)
    r = kkt_symv(N, cp, rv, nz, x) - b
    assert np.max(np.abs(r)) <= 1e-9 * np.max(np.abs(b))
    assert s.linear_solver_info().positive_inertia == 100_000
