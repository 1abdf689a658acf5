"""Pins the IPM oracle (oracle/ipm_oracle.c) against the reference's own
end-to-end known answers (tests/*.rs) and KKT-structure goldens
(kkt_assembly.rs:185-355).  CPU only."""
import numpy as np
import pytest
import scipy.sparse as sp

import oracle
import ref_problems as rp


def solve(P, q, A, b, cones, **kw):
    ipm = oracle.IPM(P, q, A, b, cones, settings=oracle.default_settings(**kw) if kw else None)
    ipm.set_perm(np.arange(ipm.N))
    return ipm, ipm.solve()


def test_qp_feasible():  # basic_qp.rs:98-117
    _, r = solve(*rp.basic_qp())
    assert r["status"] == "Solved"
    assert np.linalg.norm(r["x"] - [0.3, 0.7]) <= 1e-6
    assert abs(r["obj_val"] - 1.8800000298331538) <= 1e-6
    assert abs(r["obj_val_dual"] - 1.8800000298331538) <= 1e-6


def test_qp_singleton_cones_identical():  # basic_qp.rs:119-142 (bit-identical results)
    P, q, A, b, _ = rp.basic_qp()
    _, r1 = solve(P, q, A, b, [("nonneg", 3), ("nonneg", 3)])
    _, r2 = solve(P, q, A, b, [("nonneg", 1)] * 6)
    _, r3 = solve(P, q, A, b, [("soc", 1)] * 6)
    for r in (r2, r3):
        assert r["status"] == r1["status"] and r["obj_val"] == r1["obj_val"]
        assert np.array_equal(r["x"], r1["x"])


def test_qp_univariate():  # basic_qp.rs:80-97
    I1 = sp.identity(1, format="csc")
    _, r = solve(I1, [0.], I1, [1.], [("nonneg", 1)])
    assert r["status"] == "Solved"
    assert abs(r["x"][0]) <= 1e-6 and abs(r["obj_val"]) <= 1e-6 and abs(r["info"].cost_dual) <= 1e-6


def test_qp_dual_infeasible_ill_cond():  # basic_qp.rs:178-204
    P, c, _, _, _ = rp.basic_qp_dual_inf()
    _, r = solve(P, c, sp.csc_matrix(np.array([[1., 1.]])), [1.], [("nonneg", 1)])
    assert r["status"] == "DualInfeasible"


def test_presolve_settable_bound():  # presolve.rs:107-114
    oracle.default_infinity()
    d = oracle.get_infinity()
    assert d == 1e20
    oracle.set_infinity(1e21)
    assert oracle.get_infinity() == 1e21
    # the bound is what the presolve compares with: 5e20 is finite under 1e21 and infinite under 1e20
    n = 3
    P = sp.identity(n, format="csc"); A = (2.0 * sp.vstack([sp.identity(n), -sp.identity(n)])).tocsc()
    b = np.ones(2 * n); b[3] = 5e20
    ipm = oracle.IPM(P, [3., -2., 1.], A, b, [("nonneg", 3), ("nonneg", 3)])
    assert ipm.m_reduced == 6
    oracle.default_infinity()
    assert oracle.get_infinity() == d
    ipm = oracle.IPM(P, [3., -2., 1.], A, b, [("nonneg", 3), ("nonneg", 3)])
    assert ipm.m_reduced == 5


def test_qp_primal_infeasible():  # basic_qp.rs:144-160
    P, q, A, b, cones = rp.basic_qp()
    b = list(b); b[0] = -1.; b[3] = -1.
    _, r = solve(P, q, A, b, cones)
    assert r["status"] == "PrimalInfeasible" and np.isnan(r["obj_val"])


def test_qp_dual_infeasible():  # basic_qp.rs:162-176
    _, r = solve(*rp.basic_qp_dual_inf())
    assert r["status"] == "DualInfeasible" and np.isnan(r["obj_val"])


def test_lp_feasible():  # basic_lp.rs:32-49
    _, r = solve(*rp.basic_lp())
    assert r["status"] == "Solved"
    assert np.linalg.norm(r["x"] - [-0.5, 0.5, -0.5]) <= 1e-8
    assert abs(r["obj_val"] + 3.) <= 1e-8 and abs(r["obj_val_dual"] + 3.) <= 1e-8


def test_lp_primal_infeasible():  # basic_lp.rs:51-67
    P, q, A, b, cones = rp.basic_lp()
    b = list(b); b[0] = -1.; b[3] = -1.
    _, r = solve(P, q, A, b, cones)
    assert r["status"] == "PrimalInfeasible"


def test_lp_dual_infeasible():  # basic_lp.rs:69-85
    P, _, A, b, cones = rp.basic_lp()
    A = A.copy(); A.data[1] = 1.
    _, r = solve(P, [1., 0., 0.], A, b, cones)
    assert r["status"] == "DualInfeasible"


def test_lp_dual_infeasible_ill_cond():  # basic_lp.rs:87-104
    P, _, A, b, cones = rp.basic_lp()
    A = A.copy(); A.data[0] = np.finfo(float).eps; A.data[1] = 0.0
    _, r = solve(P, [1., 0., 0.], A, b, cones)
    assert r["status"] == "DualInfeasible"


def test_socp_feasible():  # basic_socp.rs:56-73
    _, r = solve(*rp.basic_socp())
    assert r["status"] == "Solved"
    assert np.linalg.norm(r["x"] - [-0.5, 0.435603, -0.245459]) <= 1e-4
    assert abs(r["obj_val"] + 8.4590e-01) <= 1e-4 and abs(r["obj_val_dual"] + 8.4590e-01) <= 1e-4


def test_socp_feasible_sparse():  # basic_socp.rs:75-90 (SOC(6) -> sparse expansion)
    P, q, A, b, _ = rp.basic_socp()
    ipm, r = solve(P, q, A, b, [("nonneg", 3), ("soc", 6)])
    assert ipm.N == 3 + 9 + 2
    assert r["status"] == "Solved"


def test_socp_infeasible():  # basic_socp.rs:92-108
    P, q, A, b, cones = rp.basic_socp()
    b = list(b); b[6] = -10.
    _, r = solve(P, q, A, b, cones)
    assert r["status"] == "PrimalInfeasible"


def test_eq_constrained():  # basic_eq_constrained.rs:36-83
    I3 = sp.identity(3, format="csc")
    _, r = solve(I3, [0., 0., 0.], rp.eq_A1(), [2., 0.], [("zero", 2)])
    assert r["status"] == "Solved" and np.linalg.norm(r["x"] - [0., 1., 1.]) <= 1e-6
    _, r = solve(I3, [0.] * 3, rp.eq_A2(), [1.] * 4, [("zero", 4)])
    assert r["status"] == "PrimalInfeasible"
    P = sp.csc_matrix(np.diag([0., 1., 1.]))
    P = sp.csc_matrix((np.array([0., 1., 1.]), np.array([0, 1, 2]), np.array([0, 1, 2, 3])), shape=(3, 3))
    _, r = solve(P, [1.] * 3, rp.eq_A1(), [2., 0.], [("zero", 2)])
    assert r["status"] == "DualInfeasible"


def test_unconstrained():  # basic_unconstrained.rs:6-40
    I3 = sp.identity(3, format="csc")
    _, r = solve(I3, [1., 2., -3.], sp.csc_matrix((0, 3)), [], [])
    assert r["status"] == "Solved" and np.linalg.norm(r["x"] - [-1., -2., 3.]) <= 1e-6
    _, r = solve(sp.csc_matrix((3, 3)), [1., 0., 0.], sp.csc_matrix((0, 3)), [], [])
    assert r["status"] == "DualInfeasible"


def test_hs35_and_box_qp():
    _, r = solve(*rp.hs35())
    assert r["status"] == "Solved"
    # HS35 known optimum x* = (4/3, 7/9, 4/9), f* = 1/9 - 9 (constant 9 dropped in this form)
    assert np.linalg.norm(r["x"] - [4 / 3, 7 / 9, 4 / 9]) <= 1e-6
    _, r = solve(*rp.box_qp3())
    assert r["status"] == "Solved" and np.linalg.norm(r["x"] - [-0.5, 0.5, -0.5]) <= 1e-6


def dense_from_triu(N, cp, rv, nz):
    K = np.zeros((N, N))
    for j in range(N):
        for p in range(cp[j], cp[j + 1]):
            K[rv[p], j] = nz[p]
    return K


def kkt_PA():  # kkt_assembly.rs:187-199
    P = sp.csc_matrix(np.array([[1., 2., 4.], [0., 3., 5.], [0., 0., 6.]]))
    A = sp.csc_matrix(np.array([[7., 0., 8.], [0., 9., 10.], [1., 2., 3.]] * 2))
    return P, A


def test_kkt_assembly_nncone():  # kkt_assembly.rs:201-211, 283-292
    P, A = kkt_PA()
    ipm = oracle.IPM(P, [0.] * 3, A, [0.] * 6, [("nonneg", 6)], settings=oracle.default_settings(equilibrate_enable=0))
    N, cp, rv, nz, ds = ipm.kkt()
    nz[ipm.map("Hsblocks")] = -1.
    Ku = np.array([[1., 2., 4., 7., 0., 1., 7., 0., 1.], [0., 3., 5., 0., 9., 2., 0., 9., 2.],
                   [0., 0., 6., 8., 10., 3., 8., 10., 3.]] + [[0.] * (3 + i) + [-1.] + [0.] * (5 - i) for i in range(6)])
    assert np.array_equal(dense_from_triu(N, cp, rv, nz), Ku)
    assert ds.tolist() == [1, 1, 1] + [-1] * 6
    assert np.array_equal(ipm.map("diag_full"), cp[1:] - 1)


def test_kkt_assembly_sparse_socone():  # kkt_assembly.rs:249-261, 302-347
    P, A = kkt_PA()
    ipm = oracle.IPM(P, [0.] * 3, A, [0.] * 6, [("soc", 6)], settings=oracle.default_settings(equilibrate_enable=0))
    N, cp, rv, nz, ds = ipm.kkt()
    assert N == 11
    nz[ipm.sparse_map(0, "v")] = 2.
    nz[ipm.sparse_map(0, "u")] = 3.
    nz[ipm.sparse_map(0, "D")] = 4.
    nz[ipm.map("Hsblocks")] = -1.
    Ku = np.zeros((11, 11))
    Ku[:3, :9] = [[1., 2., 4., 7., 0., 1., 7., 0., 1.], [0., 3., 5., 0., 9., 2., 0., 9., 2.],
                  [0., 0., 6., 8., 10., 3., 8., 10., 3.]]
    for i in range(6):
        Ku[3 + i, 3 + i] = -1.
        Ku[3 + i, 9] = 2.
        Ku[3 + i, 10] = 3.
    Ku[9, 9] = Ku[10, 10] = 4.
    assert np.array_equal(dense_from_triu(N, cp, rv, nz), Ku)
    assert ds.tolist() == [1, 1, 1] + [-1] * 6 + [-1, 1]   # datamaps.rs:134-136


def test_kkt_missing_P_diagonal_gets_structural_zero():  # kkt_assembly.rs:69-70,120-121
    P = sp.csc_matrix((np.array([5.]), np.array([0]), np.array([0, 0, 1, 1])), shape=(3, 3))  # only P[0,1]
    A = sp.csc_matrix(np.eye(3))
    ipm = oracle.IPM(P, [0.] * 3, A, [0.] * 3, [("nonneg", 3)], settings=oracle.default_settings(equilibrate_enable=0))
    N, cp, rv, nz, ds = ipm.kkt()
    assert all(rv[cp[j + 1] - 1] == j for j in range(N))  # diagonal is the last entry of every column
    assert len(rv) == 1 + 3 + 3 + 3


# ---- inf-bound presolve (tests/presolve.rs:29-105, data :7-27) ----
def _presolve_data():
    n = 3
    P = sp.identity(n, format="csc")
    A = (2.0 * sp.vstack([sp.identity(n), -sp.identity(n)])).tocsc()
    return P, np.array([3., -2., 1.]), A, np.ones(2 * n), [("nonneg", 3), ("nonneg", 3)]


def test_presolve_single_unbounded():  # presolve.rs:29-44
    P, c, A, b, cones = _presolve_data()
    b[3] = 1e30
    ipm, r = solve(P, c, A, b, cones)
    assert r["status"] == "Solved" and ipm.m_reduced == 5
    assert r["z"][3] == 0.0 and r["s"][3] == 1e20


def test_presolve_single_unbounded_2():  # presolve.rs:46-61
    P, c, A, b, _ = _presolve_data()
    b[4] = 1e30
    ipm, r = solve(P, c, A, b, [("zero", 2), ("nonneg", 4)])
    assert r["status"] == "Solved" and ipm.m_reduced == 5


def test_presolve_completely_redundant_cone():  # presolve.rs:63-83
    P, c, A, b, cones = _presolve_data()
    b[:3] = 1e30
    ipm, r = solve(P, c, A, b, cones)
    assert r["status"] == "Solved" and ipm.m_reduced == 3
    assert np.array_equal(r["z"][:3], np.zeros(3)) and np.array_equal(r["s"][:3], np.full(3, 1e20))
    assert np.linalg.norm(r["x"] - [-0.5, 2., -0.5]) <= 1e-6


def test_presolve_every_constraint_redundant():  # presolve.rs:85-101
    P, c, A, b, cones = _presolve_data()
    b[:] = 1e30
    ipm, r = solve(P, c, A, b, cones)
    assert r["status"] == "Solved" and ipm.m_reduced == 0
    assert np.linalg.norm(r["x"] + c) <= 1e-6


def test_presolve_disabled_keeps_the_rows():
    P, c, A, b, cones = _presolve_data()
    b[3] = 1e30
    ipm, r = solve(P, c, A, b, cones, presolve_enable=0)
    assert ipm.m_reduced == 6     # the row stays (capped at the bound, problemdata.rs:130-131); no claim on the status
