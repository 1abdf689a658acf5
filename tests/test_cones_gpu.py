"""GPU: every cone kernel (ccone_*) against the oracle's restatement of the
reference cone code on the same inputs.  The reference has no unit-level known
answers for cone numerics, so this is oracle parity."""
import numpy as np
import pytest
import scipy.sparse as sp

import clarabel_rs_b200 as cb
import oracle

pytestmark = pytest.mark.gpu

CONES = [("zero", 3), ("nonneg", 40), ("soc", 3), ("soc", 4), ("soc", 7), ("nonneg", 5), ("soc", 300), ("soc", 2)]
TOL = 1e-12


def make():
    m = sum(d for _, d in CONES)
    n = 4
    rng = np.random.default_rng(0)
    A = sp.random(m, n, density=0.5, random_state=1, format="csc") + sp.csc_matrix((np.ones(n), (np.arange(n), np.arange(n))), shape=(m, n))
    P = sp.identity(n, format="csc")
    q, b = rng.standard_normal(n), rng.standard_normal(m)
    st = dict(equilibrate_enable=0)
    dev = cb.CudaSolver(P, q, A, b, CONES, settings=cb.default_settings(**st))
    ora = oracle.IPM(P, q, A, b, CONES, settings=oracle.default_settings(**st))
    return dev, ora, m


def interior_point(rng, m):
    v = np.zeros(m)
    o = 0
    for kind, d in CONES:
        if kind == "nonneg":
            v[o:o + d] = rng.uniform(0.1, 3.0, d)
        elif kind == "soc":
            t = rng.standard_normal(d - 1)
            v[o + 1:o + d] = t
            v[o] = np.linalg.norm(t) * rng.uniform(1.05, 2.0) + 0.1
        o += d
    return v


def close(a, b, tol=TOL):
    sc = max(1.0, np.max(np.abs(b))) if np.size(b) else 1.0
    return np.max(np.abs(np.asarray(a) - np.asarray(b))) <= tol * sc if np.size(b) else True


def test_all_cone_ops_match_oracle():
    dev, ora, m = make()
    rng = np.random.default_rng(42)
    for trial in range(3):
        s, z = interior_point(rng, m), interior_point(rng, m)
        assert dev.cone_update_scaling(s, z) and ora.update_scaling(s, z)
        assert close(dev.cone_get_Hs(), ora.get_Hs())
        x = rng.standard_normal(m)
        assert close(dev.cone_mul_Hs(x), ora.mul_Hs(x))
        assert close(dev.cone_affine_ds(), ora.affine_ds())
        dz, ds = rng.standard_normal(m), rng.standard_normal(m)
        assert close(dev.cone_combined_ds_shift(dz, ds, 0.37), ora.combined_ds_shift(dz, ds, 0.37))
        assert close(dev.cone_ds_from_dz_offset(ds, z), ora.ds_from_dz_offset(ds, z), 1e-11)
        a_dev, a_ora = dev.cone_step_length(dz, ds, z, s, 1.0), ora.step_length(dz, ds, z, s, 1.0)
        assert abs(a_dev - a_ora) <= 1e-12 * max(1.0, a_ora)
        # larger alpha_max: the cone bound, not the cap, is active
        a_dev, a_ora = dev.cone_step_length(dz, ds, z, s, 1e6), ora.step_length(dz, ds, z, s, 1e6)
        assert abs(a_dev - a_ora) <= 1e-11 * max(1.0, a_ora)


def test_scaling_failure_outside_cone():
    dev, ora, m = make()
    rng = np.random.default_rng(1)
    s, z = interior_point(rng, m), interior_point(rng, m)
    o = 3 + 40            # first SOC(3)
    s[o] = 0.0            # not interior
    assert not dev.cone_update_scaling(s, z)
    assert not ora.update_scaling(s, z)


def test_identity_scaling_margins_and_shift():
    dev, ora, m = make()
    dev.cone_set_identity_scaling()
    Hs = dev.cone_get_Hs()
    # identity scaling: NN -> 1, sparse SOC diag -> [d=0.5, 1...], dense SOC -> 2ww'-J = I, zero -> 0
    o = 0
    for kind, d in CONES:
        if kind == "zero":
            assert np.all(Hs[o:o + d] == 0); o += d
        elif kind == "nonneg":
            assert np.all(Hs[o:o + d] == 1); o += d
        elif d > 4:
            assert abs(Hs[o] - 0.5) < 1e-15 and np.all(Hs[o + 1:o + d] == 1); o += d
        else:
            blk = Hs[o:o + d * (d + 1) // 2]; o += d * (d + 1) // 2
            M = np.zeros((d, d)); k = 0
            for c in range(d):
                for r in range(c + 1):
                    M[r, c] = blk[k]; k += 1
            assert np.allclose(M, np.eye(d), atol=1e-15)
    rng = np.random.default_rng(3)
    z = rng.standard_normal(m)
    mn, ps = dev.cone_margins(z)
    # reference semantics (compositecone.rs:197-205): min over cones, sum of positive parts
    exp_min, exp_pos, o = np.inf, 0.0, 0
    for kind, d in CONES:
        if kind == "nonneg":
            exp_min = min(exp_min, z[o:o + d].min()); exp_pos += np.maximum(z[o:o + d], 0).sum()
        elif kind == "soc":
            a = z[o] - np.linalg.norm(z[o + 1:o + d]); exp_min = min(exp_min, a); exp_pos += max(a, 0.0)
        o += d
    assert abs(mn - exp_min) < 1e-13 and abs(ps - exp_pos) < 1e-12
    zs = dev.cone_scaled_unit_shift(z, 2.5, True)
    o = 0
    for kind, d in CONES:
        if kind == "zero":
            assert np.all(zs[o:o + d] == 0)
        elif kind == "nonneg":
            assert np.allclose(zs[o:o + d], z[o:o + d] + 2.5)
        else:
            assert zs[o] == z[o] + 2.5 and np.array_equal(zs[o + 1:o + d], z[o + 1:o + d])
        o += d
    zs = dev.cone_scaled_unit_shift(z, 2.5, False)
    assert np.array_equal(zs[:3], z[:3])  # dual zero cone untouched (zerocone.rs:63-69)
