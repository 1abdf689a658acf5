"""GPU: the reference's tests/data_updating.rs through the C-ABI -- every update form (matrix, value vector,
(index, values) pairs) on P, A, q, b followed by a solve equals a solver built from the changed data to 1e-7, the
no-op combinations are accepted, updates are refused while the presolver has removed rows."""
import numpy as np
import pytest
import scipy.sparse as sp

import clarabel_rs_b200 as cb

pytestmark = pytest.mark.gpu


def data():      # data_updating.rs:8-44: huge values so that the equilibration scalings are far from 1 and carry through
    P = sp.csc_matrix(np.array([[40000., 1.], [1., 20000.]]))
    I2 = sp.identity(2, format="csc")
    A = sp.vstack([-I2, I2]).tocsc()
    return P, np.array([10000., 10000.]), A, np.ones(4), [("nonneg", 2), ("nonneg", 2)]


def settings(**kw):
    return cb.default_settings(presolve_enable=0, equilibrate_enable=1, **kw)


def solved(P, q, A, b, cones, st=None):
    s = cb.CudaSolver(P, q, A, b, cones, settings=st or settings())
    r = s.solve()
    assert r["status"] == "Solved"
    return s, r


def check(update, P2=None, q2=None, A2=None, b2=None):
    P, q, A, b, cones = data()
    s1, _ = solved(P, q, A, b, cones)
    s1.update_data(**update)
    r1 = s1.solve()
    _, r2 = solved(P2 if P2 is not None else P, q2 if q2 is not None else q, A2 if A2 is not None else A,
                   b2 if b2 is not None else b, cones)
    assert r1["status"] == "Solved"
    assert np.linalg.norm(r1["x"] - r2["x"]) <= 1e-7


def test_update_P_matrix_and_vector_form():      # :47-89
    P = data()[0]
    P2 = sp.triu(P, format="csc"); P2.data[0] = 100.
    check(dict(P=P2), P2=P2)
    check(dict(P=P2.data.copy()), P2=P2)


def test_update_P_tuple():      # :91-115
    P00 = 40000.
    check(dict(P=([1, 2], [3., 5.])), P2=sp.csc_matrix(np.array([[P00, 3.], [0., 5.]])))


def test_update_A_matrix_vector_and_tuple_form():      # :117-189
    A = data()[2]
    A2 = A.copy(); A2.data[2] = -1000.      # entry (1, 1) of [-I; I], as the reference's own assert says (:124-128)
    assert A2[1, 1] == -1000.
    check(dict(A=A2), A2=A2)
    check(dict(A=A2.data.copy()), A2=A2)
    A3 = A.copy(); A3.data[1] = 0.5; A3.data[2] = -0.5
    check(dict(A=([1, 2], [0.5, -0.5])), A2=A3)


def test_update_q_and_q_tuple():      # :191-234
    q = data()[1]
    q2 = q.copy(); q2[1] = 10.
    check(dict(q=q2), q2=q2)
    q3 = q.copy(); q3[1] = 10.
    check(dict(q=([1], [10.])), q2=q3)


def test_update_b_and_b_tuple():      # :236-278
    b = data()[3]
    b2 = b.copy(); b2[0] = 0.
    check(dict(b=b2), b2=b2)
    b3 = b.copy(); b3[1] = 0.; b3[3] = 0.
    check(dict(b=([1, 3], [0., 0.])), b2=b3)


def test_update_noops():      # :280-309
    P, q, A, b, cones = data()
    s, _ = solved(P, q, A, b, cones)
    for kw in [dict(P=[]), dict(A=[]), dict(q=[]), dict(b=[])]:
        s.update_data(**kw)
    P2, bz = sp.triu(P, format="csc"), ([1, 3], [0., 0.])
    for kw in [dict(P=[], q=[], A=[], b=[]), dict(P=P2, q=[], A=A, b=[]), dict(P=P2.data, q=[], A=A.data, b=[]),
               dict(P=P2, q=[], A=A.data, b=[]), dict(P=[], q=q, A=[], b=bz), dict(P=P2.data, q=[], A=A, b=bz),
               dict(P=[], q=q, A=[], b=b), dict(P=P2, q=q, A=[], b=[]), dict(P=[], q=[], A=A, b=b)]:
        s.update_data(**kw)
    assert s.solve()["status"] == "Solved"


def test_fail_on_presolve_enable():      # :311-357
    P, q, A, b, cones = data()
    st = cb.default_settings(presolve_enable=1, equilibrate_enable=1)
    assert cb.CudaSolver(P, q, A, b, cones, settings=st).is_data_update_allowed()      # enabled, nothing eliminated
    b = b.copy(); b[0] = 1e40
    assert cb.CudaSolver(P, q, A, b, cones, settings=settings()).is_data_update_allowed()      # disabled in the settings
    s = cb.CudaSolver(P, q, A, b, cones, settings=st)      # a row is eliminated
    assert not s.is_data_update_allowed()
    for kw in [dict(P=[]), dict(A=[]), dict(b=[]), dict(q=[])]:
        with pytest.raises(cb.DataUpdateError) as e:
            s.update_data(**kw)
        assert "PresolveIsActive" in str(e.value)


@pytest.mark.gpu
def test_update_with_a_different_sparsity_pattern_is_refused():
    """CscMatrix::is_equal_sparsity (algebra/csc/core.rs:436-445): a matrix with the same number of nonzeros in other
    places is an error, not a silent write into the wrong KKT slots."""
    import scipy.sparse as sp
    P = sp.csc_matrix(np.array([[4.0, 1.0], [1.0, 2.0]]))
    A = sp.csc_matrix(np.array([[1.0, 0.0], [0.0, 1.0], [1.0, 1.0]]))
    s = cb.CudaSolver(P, np.ones(2), A, np.ones(3), [("nonneg", 3)], settings=cb.default_settings(presolve_enable=0))
    s.solve()
    A2 = sp.csc_matrix(np.array([[1.0, 1.0], [0.0, 1.0], [1.0, 0.0]]))      # 4 nonzeros as well, elsewhere
    with pytest.raises(cb.DataUpdateError):
        s.update_data(A=A2)
    s.update_data(A=A * 2.0)
    assert s.solve()["status"] == "Solved"
