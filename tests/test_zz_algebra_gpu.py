"""GPU: the sparse products and reductions of the iteration body (k_csr_spmv, k_sum, k_max_nonneg) against the
reference's own unit-test answers (src/algebra/tests/matrix.rs, vector.rs) and against the oracle on random data."""
import ctypes as C

import numpy as np
import pytest
import scipy.sparse as sp

import clarabel_rs_b200 as cb
import oracle
from test_oracle_algebra import A34, TRIU

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def handle():      # P = the symmetric 4 x 4 of matrix.rs:4-15, A = the 3 x 4 of matrix.rs:64-73, nothing rescaled
    Pu = sp.csc_matrix((TRIU[2], TRIU[1], TRIU[0]), shape=(4, 4))
    A = sp.csc_matrix((A34[2], A34[1], A34[0]), shape=(3, 4))
    return cb.CudaSolver(Pu, np.zeros(4), A, np.ones(3), [("nonneg", 3)], settings=cb.default_settings(equilibrate_enable=0))


def test_symv_known_answer(handle):      # matrix.rs:250-266
    assert np.array_equal(handle.test_spmv(0, [0., 1., -1., 2.], [1., 2., -3., -4.], -2., 3.), [46.0, -29.0, -25.0, -4.0])


def test_gemv_known_answers(handle):      # matrix.rs:232-247
    assert np.array_equal(handle.test_spmv(1, [5., -6., 7.], [1., -2., 3., -4.], 2., -3.), [7., 66., 35.])
    assert np.array_equal(handle.test_spmv(2, [1., -2., 3., -4.], [5., -6., 7.], 2., -3.), [-49., -220., -33., 42.])


def test_norms_and_dot_known_answers(handle):      # vector.rs:103-180
    for x in ([-3., -4., -12.], [4., -3., 12.], [-12., 3., 4.]):
        assert handle.test_vec(0, x) == 13.0
    assert handle.test_vec(0, []) == 0.0
    for x, s in (([-3. / 2., -4. / 3., -12. / 4.], [-2., 3., 4.]), ([4. / 3., -3. / 2., 12. / 4.], [3., -2., 4.])):
        assert abs(handle.test_vec(2, x, s) - 13.0) <= 1e-14 * 13.0
    assert handle.test_vec(1, [-3., 4., -12.]) == 12.0
    assert np.isnan(handle.test_vec(1, [-3., np.nan, -12.]))      # NaN propagates (vecmath.rs:132-141)
    assert handle.test_vec(3, [3., 0., 2., 1.], [-1., -2., 3., 4.]) == 7.0


def test_products_and_reductions_match_the_oracle_on_random_data():
    rng = np.random.default_rng(0)
    n, m = 300, 500
    Pu = sp.triu(sp.random(n, n, density=0.02, random_state=1) + sp.identity(n), format="csc")
    A = sp.random(m, n, density=0.02, random_state=2, format="csc")
    Pu.sort_indices(); A.sort_indices()
    dev = cb.CudaSolver(Pu, np.zeros(n), A, np.ones(m), [("nonneg", m)], settings=cb.default_settings(equilibrate_enable=0))
    L = oracle._ipm_lib()
    L.oipm_test_symv.restype = None
    L.oipm_test_gemv.restype = None
    i64 = lambda a: np.ascontiguousarray(a, dtype=np.int64)
    pi = lambda a: a.ctypes.data_as(C.POINTER(C.c_int64))
    pf = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    x, y = rng.standard_normal(n), rng.standard_normal(n)
    yo = y.copy()
    L.oipm_test_symv(C.c_int64(n), pi(i64(Pu.indptr)), pi(i64(Pu.indices)), pf(Pu.data), pf(yo), pf(x), C.c_double(0.7), C.c_double(-1.3))
    assert np.max(np.abs(dev.test_spmv(0, y, x, 0.7, -1.3) - yo)) <= 1e-13 * max(1.0, np.max(np.abs(yo)))
    for trans, nx, ny in ((0, n, m), (1, m, n)):
        xx, yy = rng.standard_normal(nx), rng.standard_normal(ny)
        yo = yy.copy()
        L.oipm_test_gemv(C.c_int64(m), C.c_int64(n), pi(i64(A.indptr)), pi(i64(A.indices)), pf(A.data), C.c_int(trans), pf(yo), pf(xx),
                         C.c_double(-0.4), C.c_double(2.0))
        assert np.max(np.abs(dev.test_spmv(1 + trans, yy, xx, -0.4, 2.0) - yo)) <= 1e-13 * max(1.0, np.max(np.abs(yo)))
    big = rng.standard_normal(100_000)
    w = rng.uniform(0.5, 2.0, big.size)
    assert abs(dev.test_vec(0, big) - np.linalg.norm(big)) <= 1e-12 * np.linalg.norm(big)
    assert dev.test_vec(1, big) == np.max(np.abs(big))
    assert abs(dev.test_vec(2, big, w) - np.linalg.norm(big * w)) <= 1e-12 * np.linalg.norm(big * w)
    assert abs(dev.test_vec(3, big, w) - float(big @ w)) <= 1e-10 * np.linalg.norm(big) * np.linalg.norm(w)


def test_two_norm_is_overflow_safe_like_stable_norm(handle):
    """vecmath.rs:206-226: the reference's 2-norm never squares an entry unscaled.  sqrt(sum x^2) would return inf for
    entries around 1e200 and 0 for entries around 1e-200; the device norms (k_norm2, Blue's three accumulators) must
    return what the reference's stable_norm returns (here: the oracle's restatement, pinned on vector.rs)."""
    cases = [[3e200, -4e200, 12e200], [3e-200, 4e-200, -12e-200], [1e200, 1.0, 1e-200], [1e-170, 1e-170, 3.0],
             [5e153, 5e153], [2e-160] * 7, [1e308, 1e308]]
    for x in cases:
        got = handle.test_vec(0, x)
        xa = np.ascontiguousarray(x, dtype=np.float64)
        Lo = oracle._ipm_lib()
        Lo.oipm_test_vec.restype = C.c_double
        want = Lo.oipm_test_vec(C.c_int(0), xa.ctypes.data_as(C.POINTER(C.c_double)), xa.ctypes.data_as(C.POINTER(C.c_double)), C.c_int64(xa.size))
        assert np.isfinite(got) and got > 0.0, (x, got)
        assert abs(got - want) <= 4e-16 * want, (x, got, want)
    # scaled variant and a long vector that mixes all three ranges
    rng = np.random.default_rng(5)
    v = np.concatenate([rng.standard_normal(5000) * 1e180, rng.standard_normal(5000), rng.standard_normal(5000) * 1e-180])
    w = rng.uniform(0.5, 2.0, v.size)
    ref = np.max(np.abs(v * w)) * np.linalg.norm((v * w) / np.max(np.abs(v * w)))
    assert abs(handle.test_vec(2, v, w) - ref) <= 1e-13 * ref
    assert np.isnan(handle.test_vec(0, [1.0, np.nan, 2.0]))
