"""The reference's own end-to-end test problems (tests/*.rs), verbatim data."""
import numpy as np
import scipy.sparse as sp


def basic_qp():  # tests/basic_qp.rs:8-43
    P = sp.csc_matrix(np.array([[4., 1.], [1., 2.]]))
    A0 = np.array([[1., 1.], [1., 0.], [0., 1.]])
    A = sp.csc_matrix(np.vstack([-A0, A0]))
    return P, [1., 1.], A, [-1., 0., 0., 1., 0.7, 0.7], [("nonneg", 3), ("nonneg", 3)]


def basic_qp_dual_inf():  # tests/basic_qp.rs:45-75
    P = sp.csc_matrix(np.array([[1., 1.], [1., 1.]]))
    A = sp.csc_matrix(np.array([[1., 1.], [1., 0.]]))
    return P, [1., -1.], A, [1., 1.], [("nonneg", 2)]


def basic_lp():  # tests/basic_lp.rs:8-30
    P = sp.csc_matrix((3, 3))
    A = sp.csc_matrix(np.vstack([np.eye(3), -np.eye(3)]) * 2.0)
    return P, [3., -2., 1.], A, [1.] * 6, [("nonneg", 3), ("nonneg", 3)]


def basic_socp():  # tests/basic_socp.rs:8-54
    nz = [1.4652521089139698, 0.6137176286085666, -1.1527861771130112, 0.6137176286085666,
          2.219109946678485, -1.4400420548730628, -1.1527861771130112, -1.4400420548730628,
          1.6014483534926371]
    P = sp.csc_matrix(np.array(nz).reshape(3, 3).T)
    A = sp.csc_matrix(np.vstack([2 * np.eye(3), -2 * np.eye(3), np.eye(3)]))
    return (P, [0.1, -2.0, 1.0], A, [1., 1., 1., 1., 1., 1., 0., 0., 0.],
            [("nonneg", 3), ("nonneg", 3), ("soc", 3)])


def eq_A1():
    return sp.csc_matrix((np.array([1., 1., 1., -1.]), np.array([0, 1, 0, 1]), np.array([0, 0, 2, 4])), shape=(2, 3))


def eq_A2():
    return sp.csc_matrix((np.array([1., 2., 1., 1., 2., -1., 1., -1., -1., 3.]),
                          np.array([2, 3, 0, 1, 2, 3, 0, 1, 2, 3]), np.array([0, 2, 6, 10])), shape=(4, 3))


def hs35():  # examples/data/hs35.json
    P = sp.csc_matrix((np.array([4.000000000000001, 2.0000000000000004, 4.000000000000001, 2.0, 2.0]),
                       np.array([0, 0, 1, 0, 2]), np.array([0, 1, 3, 5])), shape=(3, 3))
    A = sp.csc_matrix((np.array([1.0, -1.0, 1.0, -1.0, 2.0, -1.0]), np.array([0, 1, 0, 2, 0, 3]),
                       np.array([0, 2, 4, 6])), shape=(4, 3))
    return P, [-8.0, -6.0, -4.0], A, [3.0, 0.0, 0.0, 0.0], [("nonneg", 4)]


def box_qp3():  # tests/presolve.rs:7-27 (without infinite bounds)
    P = sp.identity(3, format="csc")
    A = sp.csc_matrix(np.vstack([2 * np.eye(3), -2 * np.eye(3)]))
    return P, [3., -2., 1.], A, [1.] * 6, [("nonneg", 3), ("nonneg", 3)]
