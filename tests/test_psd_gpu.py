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
