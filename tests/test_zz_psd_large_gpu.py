"""GPU: PSD-triangle cones whose matrix dimension exceeds the 32 lanes of the warp that owns them (cones_psd.cu: rows
k, k+32, ... per lane; shared memory up to n = 56, global scratch beyond), against the oracle.  Kept in a file of its
own, late in the collection order: this path was written and checked in emulation before its first device run."""
import numpy as np
import pytest
import scipy.sparse as sp

import clarabel_rs_b200 as cb
from test_oracle_psd import smat, svec
from test_psd_gpu import both, close, interior, make, numel

pytestmark = pytest.mark.gpu


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
    assert abs(rd["iterations"] - ro["iterations"]) <= 1      # equal in emulation; one iteration of slack for the device's rounding
    assert abs(rd["info"].cost_primal - ro["info"].cost_primal) <= 1e-6 * max(1.0, abs(ro["info"].cost_primal))
    X = smat(rd["x"], n)
    assert np.allclose(np.diag(X), 1.0, atol=1e-6) and np.linalg.eigvalsh(X).min() > -1e-6


def test_psd_cone_above_the_supported_dimension_is_refused():
    n = 129
    ne = n * (n + 1) // 2
    with pytest.raises(cb.BackendError):
        cb.CudaSolver(sp.identity(ne, format="csc"), np.zeros(ne), -sp.identity(ne, format="csc"), np.zeros(ne), [("psd", n)])
