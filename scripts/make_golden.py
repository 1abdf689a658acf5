"""Writes tests/golden/*.json: the reference's own end-to-end known answers as data.

The reference is Rust and cannot run here (no cargo), so these fixtures are not outputs of a run: every problem is the
data of one of the reference's integration tests and every expected value is the number that test asserts, with the
file:line it comes from.  The problem part of each file is in the reference's JSON interchange format
(default/json.rs), so the same files load into the real solver.  Regenerate with `python scripts/make_golden.py`.
"""
import json
import os
import sys

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import importlib.util

import clarabel_rs_b200 as cb
import ref_problems as rp
import test_oracle_nonsym as ns

_spec = importlib.util.spec_from_file_location("jsonio", os.path.join(os.path.dirname(cb.pkg.__file__), "jsonio.py"))
jsonio = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(jsonio)
OUT = os.path.join(ROOT, "tests", "golden")


def write(name, prob, source, status, x=None, xtol=None, obj=None, objtol=None, settings=None, notes=None):
    P, q, A, b, cones = prob
    path = os.path.join(OUT, name + ".json")
    jsonio.save_problem(path, P, q, A, b, cones, settings=settings or {})
    d = json.load(open(path))
    d["expected"] = {"source": source, "status": status}
    if x is not None:
        d["expected"].update(x=[float(v) for v in x], x_tol=xtol)
    if obj is not None:
        d["expected"].update(obj=float(obj), obj_tol=objtol)
    if notes:
        d["expected"]["notes"] = notes
    json.dump(d, open(path, "w"))


def presolve_data():
    n = 3
    return (sp.identity(n, format="csc"), np.array([3., -2., 1.]),
            (2.0 * sp.vstack([sp.identity(n), -sp.identity(n)])).tocsc(), np.ones(2 * n), [("nonneg", 3), ("nonneg", 3)])


def main():
    os.makedirs(OUT, exist_ok=True)
    write("basic_qp", rp.basic_qp(), "tests/basic_qp.rs:98-117", "Solved", x=[0.3, 0.7], xtol=1e-6,
          obj=1.8800000298331538, objtol=1e-6)
    P, q, A, b, cones = rp.basic_qp()
    b = list(b); b[0] = -1.; b[3] = -1.
    write("basic_qp_primal_infeasible", (P, q, A, b, cones), "tests/basic_qp.rs:144-160", "PrimalInfeasible")
    write("basic_qp_dual_infeasible", rp.basic_qp_dual_inf(), "tests/basic_qp.rs:162-176", "DualInfeasible")
    write("basic_lp", rp.basic_lp(), "tests/basic_lp.rs:32-49", "Solved", x=[-0.5, 0.5, -0.5], xtol=1e-8, obj=-3.0, objtol=1e-8)
    write("basic_socp", rp.basic_socp(), "tests/basic_socp.rs:56-73", "Solved", x=[-0.5, 0.435603, -0.245459], xtol=1e-4,
          obj=-8.4590e-01, objtol=1e-4)
    write("hs35", rp.hs35(), "examples/data/hs35.json + Hock-Schittkowski problem 35", "Solved",
          x=[4 / 3, 7 / 9, 4 / 9], xtol=1e-6)
    write("basic_expcone", ns.expcone_data(), "tests/basic_expcone.rs:38-55", "Solved", x=[5.0, 1.0, float(np.exp(5.0))],
          xtol=1e-6, obj=-5.0, objtol=1e-6)
    P, c, A, b, cones = ns.expcone_data()
    b = b.copy(); b[4] = -1.
    write("basic_expcone_primal_infeasible", (P, c, A, b, cones), "tests/basic_expcone.rs:57-72", "PrimalInfeasible")
    write("basic_expcone_dual_infeasible", (sp.csc_matrix((3, 3)), [-1., 0., 0.], -sp.identity(3, format="csc"), np.zeros(3),
                                            [("exp", 3)]), "tests/basic_expcone.rs:74-91", "DualInfeasible")
    n = 6
    A = sp.vstack([-sp.identity(n, format="csc"), sp.csc_matrix(np.array([[1., 2., 0., 3., 0., 0.], [0., 0., 0., 0., 1., 0.]]))]).tocsc()
    b = np.concatenate([np.zeros(n), [3., 1.]])
    write("basic_powcone", (sp.csc_matrix((n, n)), np.array([0., 0., -1., 0., 0., -1.]), A, b,
                            [("pow", 0.6), ("pow", 0.1), ("zero", 2)]), "tests/basic_powcone.rs:5-52", "Solved",
          obj=-1.8458, objtol=1e-3)
    write("basic_genpowcone", ns.genpow_data(), "tests/basic_genpowcone.rs:5-45", "Solved", obj=-1.8458, objtol=1e-3)
    write("mixed_conic", ns.mixed_conic_data(), "tests/mixed_conic.rs:5-35", "Solved", obj=0.0, objtol=1e-8)
    write("mixed_conic_dual_strategy", ns.mixed_conic_data(), "tests/mixed_conic.rs:37-46", "Solved", obj=0.0, objtol=1e-8,
          settings={"min_switch_step_length": 0.999}, notes="min_switch_step_length = 0.999 forces the dual scaling strategy")
    P, c, A, b, cones = presolve_data()
    b = b.copy(); b[:3] = 1e30
    write("presolve_redundant_cone", (P, c, A, b, cones), "tests/presolve.rs:63-83", "Solved", x=[-0.5, 2., -0.5], xtol=1e-6,
          notes="rows 0..2 have an infinite bound: z[0:3] = 0 and s[0:3] = 1e20 in the solution")
    import test_oracle_psd as psd
    write("basic_sdp", psd.sdp_data(), "tests/basic_sdp.rs:44-58", "Solved", x=psd.REFSOL, xtol=1e-6, obj=psd.REFOBJ, objtol=1e-6)
    P, q, A, b, cones = psd.sdp_data()
    write("basic_sdp_primal_infeasible", (P, q, sp.vstack([A, -A]).tocsc(), list(b) + [0.0] * 6, cones + cones),
          "tests/basic_sdp.rs:78-95", "PrimalInfeasible")
    I3 = sp.identity(3, format="csc")
    write("basic_eq_constrained", (I3, [0., 0., 0.], rp.eq_A1(), [2., 0.], [("zero", 2)]), "tests/basic_eq_constrained.rs:35-50",
          "Solved", x=[0., 1., 1.], xtol=1e-6)
    write("basic_eq_constrained_primal_infeasible", (I3, [0.] * 3, rp.eq_A2(), [1.] * 4, [("zero", 4)]),
          "tests/basic_eq_constrained.rs:52-65", "PrimalInfeasible")
    Pd = sp.csc_matrix((np.array([0., 1., 1.]), np.array([0, 1, 2]), np.array([0, 1, 2, 3])), shape=(3, 3))
    write("basic_eq_constrained_dual_infeasible", (Pd, [1.] * 3, rp.eq_A1(), [2., 0.], [("zero", 2)]),
          "tests/basic_eq_constrained.rs:67-83", "DualInfeasible")
    write("basic_unconstrained", (I3, [1., 2., -3.], sp.csc_matrix((0, 3)), [], []), "tests/basic_unconstrained.rs:6-22", "Solved",
          x=[-1., -2., 3.], xtol=1e-6)
    write("basic_unconstrained_dual_infeasible", (sp.csc_matrix((3, 3)), [1., 0., 0.], sp.csc_matrix((0, 3)), [], []),
          "tests/basic_unconstrained.rs:24-40", "DualInfeasible")
    I1 = sp.identity(1, format="csc")
    write("basic_qp_univariate", (I1, [0.], I1, [1.], [("nonneg", 1)]), "tests/basic_qp.rs:80-97", "Solved", x=[0.], xtol=1e-6,
          obj=0.0, objtol=1e-6)
    Pd, cd, _, _, _ = rp.basic_qp_dual_inf()
    write("basic_qp_dual_infeasible_ill_cond", (Pd, cd, sp.csc_matrix(np.array([[1., 1.]])), [1.], [("nonneg", 1)]),
          "tests/basic_qp.rs:178-204", "DualInfeasible")
    print("wrote", len(os.listdir(OUT)), "files to", OUT)


if __name__ == "__main__":
    main()
