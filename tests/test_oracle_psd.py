"""Pins the oracle's PSD-triangle cone (oracle/ipm_oracle.c, psd_*) on the reference's
end-to-end SDP known answers (tests/basic_sdp.rs) and its dense kernels against numpy's LAPACK
(the reference calls dpotrf/dgesdd/dsyevr through third-party crates).  CPU only."""
import numpy as np
import pytest
import scipy.sparse as sp

import oracle


def sdp_data():  # basic_sdp.rs:8-30
    I6 = sp.identity(6, format="csc")
    return I6, [0.0] * 6, I6, [-3., 1., 4., 1., 2., 5.], [("psd", 3)]


REFSOL = [-3.0729833267361095, 0.3696004167288786, -0.022226685581313674, 0.31441213129613066,
          -0.026739700851545107, -0.016084530571308823]
REFOBJ = 4.840076866013861


def solve(P, q, A, b, cones):
    ipm = oracle.IPM(P, q, A, b, cones)
    ipm.set_perm(np.arange(ipm.N))
    return ipm, ipm.solve()


def test_sdp_feasible():  # basic_sdp.rs:47-60
    ipm, r = solve(*sdp_data())
    assert r["status"] == "Solved"
    assert np.linalg.norm(r["x"] - REFSOL) <= 1e-6
    assert abs(r["info"].cost_primal - REFOBJ) <= 1e-6


def test_sdp_empty_cone():  # basic_sdp.rs:62-78
    P, q, A, b, cones = sdp_data()
    _, r = solve(P, q, A, b, cones + [("psd", 0)])
    assert r["status"] == "Solved" and np.linalg.norm(r["x"] - REFSOL) <= 1e-6


def test_sdp_primal_infeasible():  # basic_sdp.rs:80-97
    P, q, A, b, cones = sdp_data()
    A2 = sp.vstack([A, -A]).tocsc()
    _, r = solve(P, q, A2, list(b) + [0.0] * 6, cones + cones)
    assert r["status"] == "PrimalInfeasible"


def svec(M):
    n = M.shape[0]
    out = []
    for c in range(n):
        for r in range(c + 1):
            out.append(M[r, c] if r == c else np.sqrt(2) * M[r, c])
    return np.array(out)


def smat(x, n):
    M = np.zeros((n, n)); k = 0
    for c in range(n):
        for r in range(c + 1):
            M[r, c] = M[c, r] = x[k] if r == c else x[k] / np.sqrt(2)
            k += 1
    return M


@pytest.mark.parametrize("n", [2, 5, 20])
def test_psd_scaling_against_lapack(n):
    """NT scaling identities with numpy's LAPACK as the judge: W = R-congruence, lambda = W z = W^-T s,
    Hs = W'W as a matrix on svec space, step length from the smallest eigenvalue."""
    rng = np.random.default_rng(n)
    numel = n * (n + 1) // 2
    A = sp.identity(numel, format="csc")
    ipm = oracle.IPM(sp.identity(numel, format="csc"), np.zeros(numel), A, np.zeros(numel), [("psd", n)],
                     settings=oracle.default_settings(equilibrate_enable=0))
    F, G = rng.standard_normal((n, n)), rng.standard_normal((n, n))
    S, Z = F @ F.T + 0.5 * np.eye(n), G @ G.T + 0.5 * np.eye(n)
    s, z = svec(S), svec(Z)
    assert ipm.update_scaling(s, z)
    # reference construction with LAPACK (psdtrianglecone.rs:144-204)
    L1, L2 = np.linalg.cholesky(S), np.linalg.cholesky(Z)
    U, sv, Vt = np.linalg.svd(L2.T @ L1)
    R = L1 @ Vt.T @ np.diag(sv ** -0.5)
    # Hs as an operator: y = svec(RR' X RR')
    Hs_packed = ipm.get_Hs()
    H = np.zeros((numel, numel)); k = 0
    for c in range(numel):
        for r in range(c + 1):
            H[r, c] = H[c, r] = Hs_packed[k]; k += 1
    x = rng.standard_normal(numel)
    RRt = R @ R.T
    assert np.allclose(H @ x, svec(RRt @ smat(x, n) @ RRt), rtol=1e-10, atol=1e-11)
    assert np.allclose(ipm.mul_Hs(x), H @ x, rtol=1e-10, atol=1e-11)
    # lambda^2 on the diagonal of affine_ds
    ds = ipm.affine_ds()
    lam2 = np.array([ds[k * (k + 3) // 2] for k in range(n)])
    assert np.allclose(np.sort(lam2), np.sort(sv ** 2), rtol=1e-10)
    # step length: largest alpha with Z + alpha dZ PSD (and S likewise)
    dZ, dS = smat(rng.standard_normal(numel), n), smat(rng.standard_normal(numel), n)
    a = ipm.step_length(svec(dZ), svec(dS), z, s, 1e6)

    def maxstep(X, dX):
        Li = np.linalg.inv(np.linalg.cholesky(X))
        ev = np.linalg.eigvalsh(Li @ dX @ Li.T)
        return 1e6 if ev.min() >= 0 else min(1e6, -1.0 / ev.min())
    assert abs(a - min(maxstep(Z, dZ), maxstep(S, dS))) <= 1e-9 * max(1.0, a)
