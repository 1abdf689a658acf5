"""Pins the oracle's sparse and vector kernels (symv on one stored triangle, quadratic form,
A x / A' x, overflow-safe norms, NaN-propagating infinity norm) on the reference's own unit tests:
src/algebra/tests/matrix.rs and src/algebra/tests/vector.rs.  Exact equality where the reference asserts it."""
import ctypes as C

import numpy as np

import oracle

L = oracle._ipm_lib()
I64 = lambda a: np.ascontiguousarray(a, dtype=np.int64)
F64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
pi = lambda a: a.ctypes.data_as(C.POINTER(C.c_int64))
pf = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
L.oipm_test_quad_form.restype = C.c_double
L.oipm_test_vec.restype = C.c_double
L.oipm_test_symv.restype = None
L.oipm_test_gemv.restype = None

# matrix.rs:4-15 (upper triangle) and its transpose; matrix.rs:64-73 (3 x 4)
TRIU = (I64([0, 1, 3, 6, 8]), I64([0, 0, 1, 0, 1, 2, 2, 3]), F64([4., -3., 8., 7., -1., 2., -3., 1.]))
A34 = (I64([0, 2, 4, 6, 8]), I64([0, 1, 0, 2, 0, 1, 0, 2]), F64([-1., 3., -17., -4., 6., 7., 10., -5.]))


def tril_of_triu():
    import scipy.sparse as sp
    T = sp.csc_matrix((TRIU[2], TRIU[1], TRIU[0]), shape=(4, 4)).T.tocsc()
    T.sort_indices()
    return I64(T.indptr), I64(T.indices), F64(T.data)


def symv(M, y, x, a, b):
    y = F64(y).copy()
    L.oipm_test_symv(C.c_int64(4), pi(M[0]), pi(M[1]), pf(M[2]), pf(y), pf(F64(x)), C.c_double(a), C.c_double(b))
    return y


def test_symv():      # matrix.rs:250-266: either stored triangle gives the same product
    for M in (TRIU, tril_of_triu()):
        assert np.array_equal(symv(M, [0., 1., -1., 2.], [1., 2., -3., -4.], -2., 3.), [46.0, -29.0, -25.0, -4.0])


def test_quad_form():      # matrix.rs:269-285 (the oracle keeps the upper triangle, like the KKT layer does)
    v = L.oipm_test_quad_form(C.c_int64(4), pi(TRIU[0]), pi(TRIU[1]), pf(TRIU[2]), pf(F64([0., 1., -1., 2.])), pf(F64([1., 2., -3., -4.])))
    assert v == 15.0


def gemv(trans, y, x, a, b):
    y = F64(y).copy()
    L.oipm_test_gemv(C.c_int64(3), C.c_int64(4), pi(A34[0]), pi(A34[1]), pf(A34[2]), C.c_int(trans), pf(y), pf(F64(x)), C.c_double(a),
                     C.c_double(b))
    return y


def test_gemv():      # matrix.rs:232-247
    assert np.array_equal(gemv(0, [5., -6., 7.], [1., -2., 3., -4.], 2., -3.), [7., 66., 35.])
    assert np.array_equal(gemv(1, [1., -2., 3., -4.], [5., -6., 7.], 2., -3.), [-49., -220., -33., 42.])


def vec(what, x, v=None):
    x = F64(x)
    v = F64(v) if v is not None else x
    return L.oipm_test_vec(C.c_int(what), pf(x), pf(v), C.c_int64(x.size))


def test_norms():      # vector.rs:127-180
    for x in ([-3., -4., -12.], [4., -3., 12.], [-12., 3., 4.]):
        assert vec(0, x) == 13.0
    assert vec(0, []) == 0.0
    for x, s in (([-3. / 2., -4. / 3., -12. / 4.], [-2., 3., 4.]), ([4. / 3., -3. / 2., 12. / 4.], [3., -2., 4.]),
                 ([-12. / 4., 3. / 2., 4. / 3.], [4., 2., -3.])):
        assert vec(2, x, s) == 13.0
    assert vec(2, [], []) == 0.0
    assert vec(1, [-3., 4., -12.]) == 12.0
    assert np.isnan(vec(1, [-3., np.nan, -12.]))      # NaN propagates (vecmath.rs:132-141)


def test_dot():      # vector.rs:103-110
    assert vec(3, [3., 0., 2., 1.], [-1., -2., 3., 4.]) == 7.0 and vec(3, [-1., -2., 3., 4.], [3., 0., 2., 1.]) == 7.0
