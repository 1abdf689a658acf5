"""GPU: the device interior-point solver (cipm_*) against (i) the reference's own
end-to-end known answers and (ii) the CPU oracle run on the same KKT permutation:
same status, same iteration count, same solution."""
import numpy as np
import pytest
import scipy.sparse as sp

import clarabel_rs_b200 as cb
import oracle
import ref_problems as rp
from helpers import workloads

pytestmark = pytest.mark.gpu


def both(P, q, A, b, cones, **kw):
    dev = cb.CudaSolver(P, q, A, b, cones, settings=cb.default_settings(**kw) if kw else None)
    rd = dev.solve()
    ora = oracle.IPM(P, q, A, b, cones, settings=oracle.default_settings(**kw) if kw else None)
    ora.set_perm(dev.kkt_perm())
    ro = ora.solve()
    return dev, rd, ora, ro


def assert_parity(rd, ro, xtol=1e-7):
    assert rd["status"] == ro["status"]
    assert rd["iterations"] == ro["iterations"]
    if rd["status"] == "Solved":
        sc = max(1.0, np.max(np.abs(ro["x"])))
        assert np.max(np.abs(rd["x"] - ro["x"])) <= xtol * sc
        assert abs(rd["obj_val"] - ro["obj_val"]) <= xtol * max(1.0, abs(ro["obj_val"]))


def test_qp_feasible():  # basic_qp.rs:98-117
    dev, rd, _, ro = both(*rp.basic_qp())
    assert rd["status"] == "Solved"
    assert np.linalg.norm(rd["x"] - [0.3, 0.7]) <= 1e-6
    assert abs(rd["obj_val"] - 1.8800000298331538) <= 1e-6
    assert abs(rd["obj_val_dual"] - 1.8800000298331538) <= 1e-6
    assert_parity(rd, ro)


def test_kkt_structure_identical_to_oracle():
    for prob in (rp.basic_qp(), rp.basic_socp(), (rp.basic_socp()[0], rp.basic_socp()[1], rp.basic_socp()[2],
                                                   rp.basic_socp()[3], [("nonneg", 3), ("soc", 6)])):
        dev = cb.CudaSolver(*prob)
        ora = oracle.IPM(*prob)
        N, cp, rv, nz, ds = dev.kkt()
        No, cpo, rvo, nzo, dso = ora.kkt()
        assert N == No and np.array_equal(cp, cpo) and np.array_equal(rv, rvo) and np.array_equal(ds, dso)
        assert np.allclose(nz, nzo, rtol=1e-14, atol=0)   # equilibrated data agree


def test_qp_infeasible_cases():  # basic_qp.rs:144-176
    P, q, A, b, cones = rp.basic_qp()
    b2 = list(b); b2[0] = -1.; b2[3] = -1.
    _, rd, _, ro = both(P, q, A, b2, cones)
    assert rd["status"] == "PrimalInfeasible" and np.isnan(rd["obj_val"])
    assert_parity(rd, ro)
    _, rd, _, ro = both(*rp.basic_qp_dual_inf())
    assert rd["status"] == "DualInfeasible"
    assert_parity(rd, ro)


def test_lp():  # basic_lp.rs:32-104
    _, rd, _, ro = both(*rp.basic_lp())
    assert rd["status"] == "Solved" and np.linalg.norm(rd["x"] - [-0.5, 0.5, -0.5]) <= 1e-8
    assert abs(rd["obj_val"] + 3.) <= 1e-8
    assert_parity(rd, ro)
    P, q, A, b, cones = rp.basic_lp()
    b2 = list(b); b2[0] = -1.; b2[3] = -1.
    _, rd, _, ro = both(P, q, A, b2, cones)
    assert rd["status"] == "PrimalInfeasible"
    assert_parity(rd, ro)
    A2 = A.copy(); A2.data[1] = 1.
    _, rd, _, ro = both(P, [1., 0., 0.], A2, b, cones)
    assert rd["status"] == "DualInfeasible"
    assert_parity(rd, ro)


def test_socp():  # basic_socp.rs:56-108
    _, rd, _, ro = both(*rp.basic_socp())
    assert rd["status"] == "Solved"
    assert np.linalg.norm(rd["x"] - [-0.5, 0.435603, -0.245459]) <= 1e-4
    assert abs(rd["obj_val"] + 8.4590e-01) <= 1e-4
    assert_parity(rd, ro)
    P, q, A, b, _ = rp.basic_socp()
    dev, rd, _, ro = both(P, q, A, b, [("nonneg", 3), ("soc", 6)])   # sparse expansion
    assert dev.N == 14 and rd["status"] == "Solved"
    assert_parity(rd, ro)
    b2 = list(b); b2[6] = -10.
    _, rd, _, ro = both(P, q, A, b2, [("nonneg", 3), ("nonneg", 3), ("soc", 3)])
    assert rd["status"] == "PrimalInfeasible"
    assert_parity(rd, ro)


def test_eq_constrained_and_unconstrained():  # basic_eq_constrained.rs, basic_unconstrained.rs
    I3 = sp.identity(3, format="csc")
    _, rd, _, ro = both(I3, [0., 0., 0.], rp.eq_A1(), [2., 0.], [("zero", 2)])
    assert rd["status"] == "Solved" and np.linalg.norm(rd["x"] - [0., 1., 1.]) <= 1e-6
    assert_parity(rd, ro)
    _, rd, _, ro = both(I3, [0.] * 3, rp.eq_A2(), [1.] * 4, [("zero", 4)])
    assert rd["status"] == "PrimalInfeasible"
    assert_parity(rd, ro)
    _, rd, _, ro = both(I3, [1., 2., -3.], sp.csc_matrix((0, 3)), [], [])
    assert rd["status"] == "Solved" and np.linalg.norm(rd["x"] - [-1., -2., 3.]) <= 1e-6
    assert_parity(rd, ro)


def test_hs35():
    _, rd, _, ro = both(*rp.hs35())
    assert rd["status"] == "Solved" and np.linalg.norm(rd["x"] - [4 / 3, 7 / 9, 4 / 9]) <= 1e-6
    assert_parity(rd, ro)


@pytest.mark.parametrize("n,m,window,seed", [(300, 600, 30, 1), (2000, 4000, 60, 2), (1500, 2000, None, 3)])
def test_random_sparse_qp_same_iterations(n, m, window, seed):
    pr = workloads.random_sparse_qp(n=n, m=m, nnz_per_row=4, seed=seed, window=window)
    dev, rd, ora, ro = both(pr["P"], pr["q"], pr["A"], pr["b"], pr["cones"])
    assert rd["status"] == "Solved"
    assert_parity(rd, ro, xtol=1e-6)
    # the iterates follow the same trajectory, not just the same count
    k = min(len(dev.trace), len(ora.trace))
    assert np.allclose(dev.trace[:k, 0], ora.trace[:k, 0], rtol=1e-5, atol=1e-12)   # mu
    assert np.allclose(dev.trace[1:k, 1], ora.trace[1:k, 1], rtol=1e-5)             # step lengths
    # KKT residual criterion of the north star on a live system
    info = dev.info
    assert info.n_refactor == ro["info"].n_refactor


def test_mixed_cones_socp():
    pr = workloads.portfolio_socp(n_assets=300, n_soc=12, soc_dim=11, block=50, seed=5)
    dev, rd, ora, ro = both(pr["P"], pr["q"], pr["A"], pr["b"], pr["cones"])
    assert rd["status"] == "Solved"
    assert_parity(rd, ro, xtol=1e-6)


def test_kkt_solver_trait_residual():
    """KKTSolver::update/setrhs/solve: residual of the refined solution vs the oracle's
    unregularised K (north star: 1e-9 relative on the KKT residual)."""
    pr = workloads.random_sparse_qp(n=500, m=800, nnz_per_row=4, seed=9, window=40)
    dev = cb.CudaSolver(pr["P"], pr["q"], pr["A"], pr["b"], pr["cones"])
    rng = np.random.default_rng(0)
    s, z = rng.uniform(0.5, 2.0, dev.m), rng.uniform(0.5, 2.0, dev.m)
    assert dev.cone_update_scaling(s, z)
    assert dev.kkt_update()
    rx, rz = rng.standard_normal(dev.n), rng.standard_normal(dev.m)
    dev.kkt_setrhs(rx, rz)
    ok, x, zz = dev.kkt_solve()
    assert ok
    N, cp, rv, _, _ = dev.kkt()
    nz = dev.kkt_values()
    from helpers import kkt_symv
    sol = np.concatenate([x, zz])
    r = kkt_symv(N, cp, rv, nz, sol) - np.concatenate([rx, rz])
    assert np.max(np.abs(r)) <= 1e-9 * max(1.0, np.max(np.abs(np.concatenate([rx, rz]))))
    # Hs block went in negated: diagonal of the (2,2) block equals -w^2 = -s/z
    diag = nz[cp[1:] - 1]
    assert np.allclose(diag[dev.n:dev.n + dev.m], -(s / z), rtol=1e-13)


def test_paired_solves_are_bitwise_the_unpaired_ones(monkeypatch):
    """The constant-rhs and the affine systems of an iteration share one factorisation and are independent, so
    they run concurrently on two solve contexts (KKTDevice::solve2).  Per system the arithmetic is that of the
    plain path: identical iterates, bit for bit."""
    pr = workloads.random_sparse_qp(n=3000, m=5000, nnz_per_row=4, seed=11, window=60)
    args = (pr["P"], pr["q"], pr["A"], pr["b"], pr["cones"])
    monkeypatch.delenv("CB_NO_PAIRED_SOLVES", raising=False)
    a = cb.CudaSolver(*args)
    ra = a.solve()
    monkeypatch.setenv("CB_NO_PAIRED_SOLVES", "1")
    b = cb.CudaSolver(*args, kkt_perm=a.kkt_perm())
    rb = b.solve()
    assert ra["status"] == rb["status"] == "Solved"
    assert ra["iterations"] == rb["iterations"]
    assert np.array_equal(ra["x"], rb["x"]) and np.array_equal(ra["z"], rb["z"]) and np.array_equal(ra["s"], rb["s"])
    ia, ib = a.info, b.info
    assert ia.n_ldl_solve == ib.n_ldl_solve and ia.n_refactor == ib.n_refactor


def test_update_data_then_solve_matches_fresh_solver():
    """DefaultSolver::update_data (data_updating.rs:68-163, tests/data_updating.rs): overwriting P, q, A, b in an
    existing solver and solving again gives the solution of a solver built from the new data (the handle keeps its
    symbolic analysis, plans and equilibration scalings, so the iterates differ but the optimum does not)."""
    pr = workloads.random_sparse_qp(n=400, m=700, nnz_per_row=4, seed=5, window=40)
    P, q, A, b, cones = pr["P"], pr["q"], pr["A"], pr["b"], pr["cones"]
    dev = cb.CudaSolver(P, q, A, b, cones)
    r0 = dev.solve()
    assert r0["status"] == "Solved"
    rng = np.random.default_rng(1)
    P2 = P.copy(); P2.data = P2.data * 1.3
    A2 = A.copy(); A2.data = A2.data * (1.0 + 0.05 * rng.standard_normal(A2.data.size))
    q2 = q + 0.1 * rng.standard_normal(q.size)
    b2 = b + 0.05 * np.abs(rng.standard_normal(b.size))
    dev.update_data(P=P2, q=q2, A=A2, b=b2)
    r1 = dev.solve()
    fresh = cb.CudaSolver(P2, q2, A2, b2, cones)
    rf = fresh.solve()
    ora = oracle.IPM(P2, q2, A2, b2, cones)
    ora.set_perm(fresh.kkt_perm())
    ro = ora.solve()
    assert r1["status"] == rf["status"] == ro["status"] == "Solved"
    # different equilibration scalings => different iterates; the optimum agrees to the solver's tolerances
    sc = max(1.0, np.max(np.abs(ro["x"])))
    assert np.max(np.abs(rf["x"] - ro["x"])) <= 1e-7 * sc          # same data, same permutation: same trajectory
    assert np.max(np.abs(r1["x"] - ro["x"])) <= 1e-4 * sc
    assert abs(r1["obj_val"] - ro["obj_val"]) <= 1e-6 * max(1.0, abs(ro["obj_val"]))
    # partial update: only q; everything else keeps its current (updated) value
    q3 = q2 * 0.5
    dev.update_data(q=q3)
    r2 = dev.solve()
    ora3 = oracle.IPM(P2, q3, A2, b2, cones)
    ora3.set_perm(fresh.kkt_perm())
    ro3 = ora3.solve()
    assert r2["status"] == ro3["status"] == "Solved"
    assert np.max(np.abs(r2["x"] - ro3["x"])) <= 1e-4 * max(1.0, np.max(np.abs(ro3["x"])))
    assert abs(r2["obj_val"] - ro3["obj_val"]) <= 1e-6 * max(1.0, abs(ro3["obj_val"]))
