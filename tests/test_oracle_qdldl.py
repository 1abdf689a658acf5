"""Pins the CPU oracle (oracle/qdldl_oracle.c) against every known-answer test
the reference holds for its LDL path: /root/reference/src/qdldl/test.rs and
the trait-level golden of .../ldlsolvers/faer_ldl.rs:352-409.  CPU only."""
import ctypes as C

import numpy as np
import pytest

import oracle
from oracle import F, I, P, QDLDL, QDLDLError


def matrix_4x4():
    # qdldl/test.rs:5-21
    return (4, 4), [0, 1, 3, 6, 8], [0, 0, 1, 0, 1, 2, 2, 3], [8., -3., 8., 2., -1., 8., -1., 1.]


def test_invperm():  # test.rs:30-46
    L = oracle.lib()
    out = np.zeros(4, dtype=np.int64)
    assert L.oq_invperm(4, P(I([3, 0, 2, 1])), P(out)) == 0
    assert L.oq_invperm(4, P(I([3, 0, 2, 0])), P(out)) != 0  # repeated index
    assert L.oq_invperm(4, P(I([4, 0, 2, 1])), P(out)) != 0  # index too big


def test_permute():  # test.rs:48-61
    L = oracle.lib()
    perm, b = I([3, 0, 2, 1]), F([1., 2., 3., 4.])
    x, y = np.zeros(4), np.zeros(4)
    L.oq_permute(4, P(x), P(b), P(perm))
    assert x.tolist() == [4., 1., 3., 2.]
    L.oq_ipermute(4, P(y), P(x), P(perm))
    assert y.tolist() == b.tolist()


def test_solve_from_factors_exact():  # test.rs:63-100, exact equality
    L = oracle.lib()
    Lp, Li = I([0, 2, 4, 5, 5]), I([1, 2, 2, 3, 3])
    Lx, dinv = F([1., 2., 1., 7., -3.]), F([0.25, -1.0, -0.5, 1.0])
    x = [-3., 2., 1., 4.]
    b = F([-3., -1., -3., 15.])
    L.oq_lsolve(4, P(Lp), P(Li), P(Lx), P(b))
    assert b.tolist() == x
    b = F([1., 31., -11., 4.])
    L.oq_ltsolve(4, P(Lp), P(Li), P(Lx), P(b))
    assert b.tolist() == x
    b = F([4., -27., -1., -279.])
    L.oq_solve_factors(4, P(Lp), P(Li), P(Lx), P(dinv), P(b))
    assert b.tolist() == x


def test_etree():  # test.rs:102-121
    L = oracle.lib()
    _, Ap, Ai, _ = matrix_4x4()
    work, Lnz, et = np.zeros(12, np.int64), np.zeros(4, np.int64), np.zeros(4, np.int64)
    L.oq_etree(4, P(I(Ap)), P(I(Ai)), P(work), P(Lnz), P(et))
    assert et.tolist() == [1, 2, 3, -1]  # -1 stands for QDLDL_UNKNOWN (usize::MAX)


def test_permute_symmetric():  # test.rs:131-164
    L = oracle.lib()
    _, Ap, Ai, Ax = matrix_4x4()
    Ap, Ai, Ax = I(Ap), I(Ai), F(Ax)
    Pc, Pr, Pv, mp = np.zeros(5, np.int64), np.zeros(8, np.int64), np.zeros(8), np.zeros(8, np.int64)
    L.oq_permute_symmetric(4, P(Ap), P(Ai), P(Ax), P(I([0, 1, 2, 3])), P(Pc), P(Pr), P(Pv), P(mp))
    assert Pc.tolist() == Ap.tolist() and Pr.tolist() == Ai.tolist() and Pv.tolist() == Ax.tolist()
    assert mp.tolist() == list(range(8))
    Ax2 = F(np.arange(1, 9))
    perm = I([2, 3, 0, 1])
    iperm = np.zeros(4, np.int64)
    assert L.oq_invperm(4, P(perm), P(iperm)) == 0
    L.oq_permute_symmetric(4, P(Ap), P(Ai), P(Ax2), P(iperm), P(Pc), P(Pr), P(Pv), P(mp))
    assert Pc.tolist() == [0, 1, 3, 5, 8]
    assert Pr.tolist() == [0, 0, 1, 2, 0, 2, 3, 0]  # unsorted columns, as the reference notes
    assert Pv.tolist() == [6.0, 7.0, 8.0, 1.0, 4.0, 2.0, 3.0, 5.0]


@pytest.mark.parametrize("perm", [[0, 1, 2, 3], [3, 0, 1, 2], [3, 0, 2, 1]])
def test_solve_basic(perm):  # test.rs:194-230 ([3,0,1,2] is the AMD result pinned at :124-129)
    shape, Ap, Ai, Ax = matrix_4x4()
    f = QDLDL(shape, Ap, Ai, Ax, perm)
    x = f.solve([20.0, -22.0, 32.0, -7.0])
    assert np.max(np.abs(x - [1., -2., 3., -4.])) <= 1e-8


def test_solve_logical_panics():  # test.rs:232-247
    shape, Ap, Ai, Ax = matrix_4x4()
    f = QDLDL(shape, Ap, Ai, Ax, [3, 0, 1, 2], logical=True)
    with pytest.raises(AssertionError):
        f.solve([20.0, -22.0, 32.0, -7.0])


def test_solve_logical_refactor():  # test.rs:249-264
    shape, Ap, Ai, Ax = matrix_4x4()
    f = QDLDL(shape, Ap, Ai, Ax, [3, 0, 1, 2], logical=True)
    f.refactor()
    x = f.solve([20.0, -22.0, 32.0, -7.0])
    assert np.max(np.abs(x - [1., -2., 3., -4.])) <= 1e-8


def test_bad_numeric_pivot():  # test.rs:266-283
    shape, Ap, Ai, Ax = matrix_4x4()
    # the reference runs this under its AMD ordering, pinned to [3,0,1,2] at test.rs:124-129
    for perm in ([0, 1, 2, 3], [3, 0, 1, 2]):
        A0 = list(Ax); A0[0] = 0.0
        with pytest.raises(QDLDLError) as e:
            QDLDL(shape, Ap, Ai, A0, perm, regularize_enable=False)
        assert e.value.code == -4
    A1 = list(Ax); A1[-1] = 0.0
    with pytest.raises(QDLDLError) as e:
        QDLDL(shape, Ap, Ai, A1, [3, 0, 1, 2], regularize_enable=False)
    assert e.value.code == -4


def test_lower_triangular_and_zero_column():  # test.rs:285-318
    # dense 3x3 (has entries below the diagonal)
    with pytest.raises(QDLDLError) as e:
        QDLDL((3, 3), [0, 3, 6, 9], [0, 1, 2] * 3, [1., 2., 1., 3., 3., 4., 5., 6., 7.], [0, 1, 2], logical=True)
    assert e.value.code == -3
    # middle column empty; note col 0 has a sub-diagonal entry too, so the
    # reference reports NotUpperTriangular first -- either way an error
    with pytest.raises(QDLDLError):
        QDLDL((3, 3), [0, 2, 2, 5], [0, 2, 0, 1, 2], [1., 1., 5., 6., 7.], [0, 1, 2], logical=True)
    with pytest.raises(QDLDLError) as e:
        QDLDL((3, 3), [0, 1, 1, 3], [0, 0, 2], [1., 5., 7.], [0, 1, 2], logical=True)
    assert e.value.code == -2


def kkt_6x6():
    # faer_ldl.rs:353-361
    return ((6, 6), [0, 1, 2, 4, 6, 8, 10], [0, 1, 0, 2, 1, 3, 0, 4, 1, 5],
            [1.0, 2.0, 1.0, -1.0, 1.0, -2.0, -1.0, -3.0, -1.0, -4.0], [1, 1, -1, -1, -1, -1])


@pytest.mark.parametrize("perm", [list(range(6)), [5, 4, 3, 2, 1, 0], [2, 0, 4, 1, 5, 3]])
def test_trait_level_golden(perm):  # faer_ldl.rs:352-409 (backend-agnostic KAT)
    shape, Ap, Ai, Ax, ds = kkt_6x6()
    f = QDLDL(shape, Ap, Ai, Ax, perm, dsigns=ds, logical=True,
              regularize_eps=1e-13, regularize_delta=2e-7)
    # map property: KKT.nzval[i] == perm_nzval[AtoPAPt[i]]
    _, _, pv = f.permA
    assert all(Ax[i] == pv[f.AtoPAPt[i]] for i in range(len(Ax)))
    assert f.refactor_ok()
    b = [1.0, 2.0, 3.0, 4., 5., 6.]
    x = f.solve(b)
    xsol = [1.0, 0.9090909090909091, -2.0, -1.5454545454545454, -2.0, -1.7272727272727275]
    assert np.max(np.abs(x - xsol)) < 1e-10
    f.update_values([9], [-10.0])
    assert f.refactor_ok()
    x = f.solve(b)
    xsol = [1.0, 1.3076923076923077, -2.0, -1.346153846153846, -2.0, -0.7307692307692306]
    assert np.max(np.abs(x - xsol)) < 1e-10
    f.offset_values([1, 2], 3., [1, -1])
    f.scale_values([1, 2], 2.)
    _, _, pv = f.permA
    assert pv[f.AtoPAPt[1]] == (2.0 + 3.0) * 2.0 and pv[f.AtoPAPt[2]] == (1.0 - 3.0) * 2.0


def test_regularisation_rule():
    # qdldl.rs:645-651: D[k]*sign < eps  =>  D[k] = delta*sign, counted
    f = QDLDL((2, 2), [0, 1, 2], [0, 1], [1e-14, 5.0], [0, 1], dsigns=[1, -1],
              regularize_eps=1e-12, regularize_delta=1e-7)
    assert f.regularize_count == 2
    assert f.D.tolist() == [1e-7, -1e-7]
    assert f.positive_inertia == 1
