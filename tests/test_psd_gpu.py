"""GPU: PSD-triangle cone kernels and SDP solves against the oracle and the reference's
tests/basic_sdp.rs known answers."""
import numpy as np
import pytest
import scipy.sparse as sp

import clarabel_rs_b200 as cb
import oracle
from helpers import workloads
from test_oracle_psd import REFOBJ, REFSOL, sdp_data, smat, svec

pytestmark = pytest.mark.gpu

CONES = [("nonneg", 7), ("psd", 3), ("soc", 5), ("psd", 1), ("psd", 12), ("zero", 2), ("psd", 20), ("psd", 2)]


def numel(kind, d):
    return d * (d + 1) // 2 if kind == "psd" else d


def make(CONES=CONES):
    m = sum(numel(k, d) for k, d in CONES)
    n = 5
    rng = np.random.default_rng(0)
    A = sp.random(m, n, density=0.3, random_state=1, format="csc")
    P = sp.identity(n, format="csc")
    st = dict(equilibrate_enable=0)
    dev = cb.CudaSolver(P, rng.standard_normal(n), A, rng.standard_normal(m), CONES, settings=cb.default_settings(**st))
    ora = oracle.IPM(P, rng.standard_normal(n), A, rng.standard_normal(m), CONES, settings=oracle.default_settings(**st))
    return dev, ora, m


def interior(rng, CONES=CONES):
    out = []
    for kind, d in CONES:
        if kind == "zero":
            out.append(np.zeros(d))
        elif kind == "nonneg":
            out.append(rng.uniform(0.1, 3.0, d))
        elif kind == "soc":
            t = rng.standard_normal(d - 1)
            out.append(np.concatenate([[np.linalg.norm(t) * 1.3 + 0.1], t]))
        else:
            F = rng.standard_normal((d, d))
            out.append(svec(F @ F.T + 0.3 * np.eye(d)))
    return np.concatenate(out)


def close(a, b, tol):
    sc = max(1.0, np.max(np.abs(b)))
    return np.max(np.abs(np.asarray(a) - np.asarray(b))) <= tol * sc


def test_psd_cone_ops_match_oracle():
    dev, ora, m = make()
    # note: the collapsing step turns PSD(1) into a nonnegative cone on both sides (supportedcone.rs:105-161)
    rng = np.random.default_rng(5)
    for trial in range(2):
        s, z = interior(rng), interior(rng)
        assert dev.cone_update_scaling(s, z) and ora.update_scaling(s, z)
        assert close(dev.cone_get_Hs(), ora.get_Hs(), 1e-10)
        x = rng.standard_normal(m)
        assert close(dev.cone_mul_Hs(x), ora.mul_Hs(x), 1e-10)
        assert close(np.sort(dev.cone_affine_ds()), np.sort(ora.affine_ds()), 1e-10)   # lambda order-free check
        dz, ds = rng.standard_normal(m), rng.standard_normal(m)
        a_dev, a_ora = dev.cone_step_length(dz, ds, z, s, 1e6), ora.step_length(dz, ds, z, s, 1e6)
        assert abs(a_dev - a_ora) <= 1e-9 * max(1.0, a_ora)
        mn, ps = dev.cone_margins(z)
        assert mn > 0
    # identity scaling: Hs blocks are identities, margins of the identity element are 1
    dev.cone_set_identity_scaling()
    e = np.zeros(m)
    e = dev.cone_scaled_unit_shift(e, 1.0, False)
    mn, ps = dev.cone_margins(e + np.where(np.array(sum(([k == "zero"] * numel(k, d) for k, d in CONES), [])), 0.0, 0.0))
    x = np.random.default_rng(1).standard_normal(m)
    y = dev.cone_mul_Hs(x)
    o = 0
    for kind, d in CONES:
        ne = numel(kind, d)
        if kind in ("psd", "nonneg", "soc"):
            assert np.allclose(y[o:o + ne], x[o:o + ne], atol=1e-13)
        o += ne


# matrix dimensions above one warp's 32 lanes: (40, 5) still works in shared memory with two cones per CTA,
# (33, 44, 60) takes the global scratch arena (psd_prepare in cones_psd.cu)
@pytest.mark.parametrize("cones", [[("psd", 40), ("nonneg", 3), ("psd", 5)], [("psd", 33), ("soc", 4), ("psd", 44), ("psd", 60)]],
                         ids=["shared-2-per-cta", "global-scratch"])
def test_large_psd_cone_ops_match_oracle(cones):
    dev, ora, m = make(cones)
    rng = np.random.default_rng(11)
    s, z = interior(rng, cones), interior(rng, cones)
    assert dev.cone_update_scaling(s, z) and ora.update_scaling(s, z)
    assert close(dev.cone_get_Hs(), ora.get_Hs(), 1e-9)
    x = rng.standard_normal(m)
    assert close(dev.cone_mul_Hs(x), ora.mul_Hs(x), 1e-9)
    assert close(np.sort(dev.cone_affine_ds()), np.sort(ora.affine_ds()), 1e-9)
    dz, ds = rng.standard_normal(m), rng.standard_normal(m)
    a_dev, a_ora = dev.cone_step_length(dz, ds, z, s, 1e6), ora.step_length(dz, ds, z, s, 1e6)
    assert abs(a_dev - a_ora) <= 1e-8 * max(1.0, a_ora)
    mn, ps = dev.cone_margins(z)
    o = 0
    ev_min, ev_pos = np.inf, 0.0
    for kind, d in cones:
        ne = numel(kind, d)
        if kind == "psd":
            ev = np.linalg.eigvalsh(smat(z[o:o + ne], d))
            ev_min, ev_pos = min(ev_min, ev.min()), ev_pos + ev[ev > 0].sum()
        elif kind == "nonneg":
            ev_min, ev_pos = min(ev_min, z[o:o + ne].min()), ev_pos + z[o:o + ne].clip(0).sum()
        elif kind == "soc":
            r = z[o] - np.linalg.norm(z[o + 1:o + ne])
            ev_min, ev_pos = min(ev_min, r), ev_pos + max(r, 0.0)
        o += ne
    assert abs(mn - ev_min) <= 1e-9 * max(1.0, abs(ev_min)) and abs(ps - ev_pos) <= 1e-9 * ev_pos
    # not positive definite -> the scaling update reports failure, as for the small cones
    d0 = cones[0][1]
    s[:numel("psd", d0)] = svec(np.diag([1.0] * (d0 - 1) + [-1.0]))
    assert not dev.cone_update_scaling(s, z)


def test_sdp_with_a_40x40_cone_same_iterations():
    """min <C, X> s.t. diag(X) = 1, X PSD (the max-cut relaxation) with X 40 x 40: one cone beyond the 32 lanes of
    the warp that owns it"""
    n = 40
    rng = np.random.default_rng(3)
    W = rng.standard_normal((n, n)); W = (W + W.T) / 2
    ne = n * (n + 1) // 2
    # variables x = svec(X); cone rows: -x + s = 0 with s in PSD(n); equality rows: X_ii = 1
    A = sp.vstack([sp.csc_matrix((np.ones(n), (np.arange(n), [k * (k + 3) // 2 for k in range(n)])), shape=(n, ne)),
                   -sp.identity(ne, format="csc")]).tocsc()
    b = np.concatenate([np.ones(n), np.zeros(ne)])
    q = svec(W)
    P = sp.csc_matrix((ne, ne))
    dev, rd, ora, ro = both(P, q, A, b, [("zero", n), ("psd", n)])
    assert rd["status"] == ro["status"] == "Solved"
    assert rd["iterations"] == ro["iterations"]
    assert abs(rd["info"].cost_primal - ro["info"].cost_primal) <= 1e-6 * max(1.0, abs(ro["info"].cost_primal))
    X = smat(rd["x"], n)
    assert np.allclose(np.diag(X), 1.0, atol=1e-6) and np.linalg.eigvalsh(X).min() > -1e-6


def test_psd_cone_above_the_supported_dimension_is_refused():
    n = 129
    ne = n * (n + 1) // 2
    with pytest.raises(cb.BackendError):
        cb.CudaSolver(sp.identity(ne, format="csc"), np.zeros(ne), -sp.identity(ne, format="csc"), np.zeros(ne), [("psd", n)])


def test_scaling_failure_not_pd():
    dev, ora, m = make()
    rng = np.random.default_rng(2)
    s, z = interior(rng), interior(rng)
    o = 7            # PSD(3) block: make S indefinite
    s[o:o + 6] = svec(np.diag([1.0, -1.0, 1.0]))
    assert not dev.cone_update_scaling(s, z)
    assert not ora.update_scaling(s, z)


def both(P, q, A, b, cones):
    dev = cb.CudaSolver(P, q, A, b, cones)
    rd = dev.solve()
    ora = oracle.IPM(P, q, A, b, cones)
    ora.set_perm(dev.kkt_perm())
    return dev, rd, ora, ora.solve()


def test_basic_sdp():  # basic_sdp.rs:47-97
    dev, rd, _, ro = both(*sdp_data())
    assert rd["status"] == "Solved"
    assert np.linalg.norm(rd["x"] - REFSOL) <= 1e-6
    assert abs(rd["info"].cost_primal - REFOBJ) <= 1e-6
    assert rd["iterations"] == ro["iterations"]
    P, q, A, b, cones = sdp_data()
    _, rd, _, ro = both(P, q, A, b, cones + [("psd", 0)])
    assert rd["status"] == "Solved" and np.linalg.norm(rd["x"] - REFSOL) <= 1e-6
    A2 = sp.vstack([A, -A]).tocsc()
    _, rd, _, ro = both(P, q, A2, list(b) + [0.0] * 6, cones + cones)
    assert rd["status"] == ro["status"] == "PrimalInfeasible"


def test_block_sdp_same_iterations():
    pr = workloads.block_sdp(n=300, n_psd=8, psd_dim=6, nnz_per_row=4, window=60, n_nonneg=20, seed=4)
    dev, rd, ora, ro = both(pr["P"], pr["q"], pr["A"], pr["b"], pr["cones"])
    assert rd["status"] == ro["status"] == "Solved"
    assert rd["iterations"] == ro["iterations"]
    assert np.max(np.abs(rd["x"] - ro["x"])) <= 1e-6 * max(1.0, np.max(np.abs(ro["x"])))
