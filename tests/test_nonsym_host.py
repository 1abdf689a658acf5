"""The per-cone thread bodies of the CUDA exponential / power cone kernels (clarabel.rs_b200/csrc/cones_nonsym.cuh),
compiled for the host by tests/host_harness/ns3_host.cpp, against the oracle (oracle/nonsym_oracle.h).  This is the
no-GPU half of the parity proof for those kernels: the same functions, the same structure-of-arrays indexing, only
the launch geometry is missing.  The GPU half is tests/test_zz_nonsym_gpu.py.  CPU only."""
import ctypes as C
import os

import numpy as np
import pytest
import scipy.sparse as sp

import oracle

_HERE = os.path.dirname(os.path.abspath(__file__))
f64p = C.POINTER(C.c_double)


def _lib():
    path = os.path.join(_HERE, "host_harness", "libns3_host.so")
    if not os.path.exists(path):
        pytest.fail("tests/host_harness/libns3_host.so missing: run `make`")
    L = C.CDLL(path)
    L.ns3h_new.restype = C.c_void_p
    L.ns3h_new.argtypes = [C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int64), f64p]
    L.ns3h_free.argtypes = [C.c_void_p]
    for nm in ["ns3h_m", "ns3h_nHs", "ns3h_n"]:
        getattr(L, nm).argtypes = [C.c_void_p]
    L.ns3h_unit_init.argtypes = [C.c_void_p, f64p, f64p]
    L.ns3h_update_scaling.argtypes = [C.c_void_p, f64p, f64p, C.c_double, C.c_int]
    L.ns3h_get_Hs.argtypes = [C.c_void_p, f64p, C.c_double]
    L.ns3h_mul_Hs.argtypes = [C.c_void_p, f64p, f64p]
    L.ns3h_copy_rows.argtypes = [C.c_void_p, f64p, f64p]
    L.ns3h_combined_shift.argtypes = [C.c_void_p, f64p, f64p, f64p, C.c_double]
    L.ns3h_step_length.argtypes = [C.c_void_p, f64p, f64p, f64p, f64p, C.c_double, C.c_double, C.c_double]
    L.ns3h_step_length.restype = C.c_double
    L.ns3h_barrier.argtypes = [C.c_void_p, f64p, f64p, f64p, f64p, C.c_double]
    L.ns3h_barrier.restype = C.c_double
    L.ns3h_wright_omega.argtypes = [C.c_double]
    L.ns3h_wright_omega.restype = C.c_double
    return L


def P(a):
    return a.ctypes.data_as(f64p)


CODES = {"zero": 0, "nonneg": 1, "soc": 2, "psd": 3, "exp": 4, "pow": 5}


class Pair:
    """the same composite cone in the host harness and in the oracle"""

    def __init__(self, cones):
        self.L = _lib()
        self.cones = cones
        ct = np.array([CODES[k] for k, _ in cones], dtype=np.int32)
        cd = np.array([3 if k in ("exp", "pow") else int(d) for k, d in cones], dtype=np.int64)
        cp = np.array([float(d) if k == "pow" else 0.0 for k, d in cones])
        self.h = self.L.ns3h_new(len(cones), ct.ctypes.data_as(C.POINTER(C.c_int32)),
                                 cd.ctypes.data_as(C.POINTER(C.c_int64)), P(cp))
        self.m = self.L.ns3h_m(self.h)
        m = self.m
        self.ora = oracle.IPM(sp.csc_matrix((m, m)), np.zeros(m), -sp.identity(m, format="csc"), np.zeros(m), cones)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ns3h_free(self.h)

    def interior_point(self, rng, spread):
        """(z, s) near the unit initialisation, inside the cones"""
        z0, s0 = self.ora.unit_initialization()
        for _ in range(100):
            z = z0 + spread * rng.standard_normal(self.m)
            s = s0 + spread * rng.standard_normal(self.m)
            zero = np.zeros(self.m)
            if np.isfinite(self.ora.compute_barrier(z, s, zero, zero, 0.0)):
                return z, s
        raise RuntimeError("no interior point found")


NS_CONES = [("exp", 3), ("pow", 0.6), ("pow", 0.1), ("exp", 3), ("pow", 0.5), ("pow", 0.93), ("exp", 3)]
MIXED = [("zero", 2), ("exp", 3), ("nonneg", 4), ("pow", 0.3), ("soc", 3), ("soc", 7), ("exp", 3), ("psd", 3), ("pow", 0.75)]


def test_wright_omega_matches_oracle():
    L, Lo = _lib(), oracle._ipm_lib()
    for z in np.concatenate([np.linspace(0.0, 6.0, 61), 10.0 ** np.arange(1, 10)]):
        a, b = L.ns3h_wright_omega(float(z)), Lo.oipm_test_wright_omega(float(z))
        assert abs(a - b) <= 4e-15 * abs(b)


@pytest.mark.parametrize("cones", [NS_CONES, MIXED], ids=["nonsymmetric", "mixed"])
def test_unit_initialization(cones):
    pr = Pair(cones)
    z, s = np.zeros(pr.m), np.zeros(pr.m)
    pr.L.ns3h_unit_init(pr.h, P(z), P(s))
    zo, so = pr.ora.unit_initialization()
    for (kind, _), o in zip(cones, np.cumsum([0] + [3 if k in ("exp", "pow") else (d * (d + 1) // 2 if k == "psd" else d) for k, d in cones])[:-1]):
        if kind in ("exp", "pow"):
            assert np.array_equal(z[o:o + 3], zo[o:o + 3]) and np.array_equal(s[o:o + 3], so[o:o + 3])


@pytest.mark.parametrize("strategy", [0, 1], ids=["primal-dual", "dual"])
@pytest.mark.parametrize("cones", [NS_CONES, MIXED], ids=["nonsymmetric", "mixed"])
def test_scaling_Hs_products_and_shift(cones, strategy):
    pr = Pair(cones)
    rng = np.random.default_rng(11 + strategy)
    nsrows = np.zeros(pr.m, dtype=bool)
    o = 0
    blocks = []
    bo = 0
    for kind, d in cones:
        rows = 3 if kind in ("exp", "pow") else (d * (d + 1) // 2 if kind == "psd" else d)
        diag = kind in ("zero", "nonneg") or (kind == "soc" and rows > 4)
        bl = rows if diag else rows * (rows + 1) // 2
        if kind in ("exp", "pow"):
            nsrows[o:o + 3] = True
            blocks.append((bo, bl))
        o += rows
        bo += bl
    for trial in range(6):
        z, s = pr.interior_point(rng, 0.05 + 0.05 * trial)
        mu = float(s @ z) / 7.0
        assert pr.ora.update_scaling_ex(s, z, mu, strategy)
        pr.L.ns3h_update_scaling(pr.h, P(s), P(z), mu, strategy)
        # Hs blocks
        Hs = np.zeros(pr.L.ns3h_nHs(pr.h))
        pr.L.ns3h_get_Hs(pr.h, P(Hs), 1.0)
        Ho = pr.ora.get_Hs()
        for b0, bl in blocks:
            assert np.allclose(Hs[b0:b0 + bl], Ho[b0:b0 + bl], rtol=1e-10, atol=1e-12)
        Hn = np.zeros_like(Hs)
        pr.L.ns3h_get_Hs(pr.h, P(Hn), -1.0)
        assert np.array_equal(Hn, -Hs)
        # y = Hs x on the nonsymmetric rows
        x = rng.standard_normal(pr.m)
        y = np.zeros(pr.m)
        pr.L.ns3h_mul_Hs(pr.h, P(y), P(x))
        yo = pr.ora.mul_Hs(x)
        assert np.allclose(y[nsrows], yo[nsrows], rtol=1e-10, atol=1e-12)
        assert np.all(y[~nsrows] == 0.0)          # other cones' rows are not touched
        # affine_ds / ds_from_dz_offset: copies
        out = np.zeros(pr.m)
        pr.L.ns3h_copy_rows(pr.h, P(out), P(s))
        assert np.array_equal(out[nsrows], pr.ora.affine_ds_ex(s)[nsrows])
        ds = rng.standard_normal(pr.m)
        out = np.zeros(pr.m)
        pr.L.ns3h_copy_rows(pr.h, P(out), P(ds))
        assert np.array_equal(out[nsrows], pr.ora.ds_from_dz_offset(ds, z)[nsrows])
        # combined_ds_shift with the third-order correction
        step_z, step_s = 0.3 * rng.standard_normal(pr.m), 0.3 * rng.standard_normal(pr.m)
        sh = np.zeros(pr.m)
        pr.L.ns3h_combined_shift(pr.h, P(sh), P(step_z.copy()), P(step_s.copy()), 0.37 * mu)
        sho = pr.ora.combined_ds_shift(step_z, step_s, 0.37 * mu)
        assert np.allclose(sh[nsrows], sho[nsrows], rtol=1e-9, atol=1e-11)


@pytest.mark.parametrize("cones", [NS_CONES, MIXED], ids=["nonsymmetric", "mixed"])
def test_step_length_is_the_sequential_composite_rule(cones):
    """max of independent per-cone backtracking counts == the reference's running alpha threaded through the cones"""
    pr = Pair(cones)
    rng = np.random.default_rng(23)
    nsrows = np.zeros(pr.m, dtype=bool)
    o = 0
    for kind, d in cones:
        rows = 3 if kind in ("exp", "pow") else (d * (d + 1) // 2 if kind == "psd" else d)
        if kind in ("exp", "pow"):
            nsrows[o:o + 3] = True
        o += rows
    seen = set()
    for trial in range(40):
        z, s = pr.interior_point(rng, 0.1)
        scale = [0.3, 1.0, 3.0, 10.0][trial % 4]
        dz, ds = scale * rng.standard_normal(pr.m), scale * rng.standard_normal(pr.m)
        # keep the symmetric cones out of the way: no movement on their rows, so the oracle's composite rule
        # reduces to ceil + nonsymmetric backtracking from alpha_max
        dz[~nsrows] = 0.0; ds[~nsrows] = 0.0
        for amax in (1.0, 0.61):
            a = pr.L.ns3h_step_length(pr.h, P(dz), P(ds), P(z), P(s), amax, 1e-4, 0.8)
            ao = pr.ora.step_length(dz, ds, z, s, amax)
            assert a == ao, (trial, amax, a, ao)
            seen.add(a)
    assert len(seen) > 4 and 0.0 in seen or len(seen) > 6


@pytest.mark.parametrize("cones", [NS_CONES, MIXED], ids=["nonsymmetric", "mixed"])
def test_barrier(cones):
    pr = Pair(cones)
    rng = np.random.default_rng(31)
    for trial in range(10):
        z, s = pr.interior_point(rng, 0.1)
        dz, ds = 0.05 * rng.standard_normal(pr.m), 0.05 * rng.standard_normal(pr.m)
        for al in (0.0, 0.5, 0.99):
            b = pr.L.ns3h_barrier(pr.h, P(z), P(s), P(dz), P(ds), al)
            bo = pr.ora.compute_barrier(z, s, dz, ds, al)
            assert np.isfinite(bo)
            assert abs(b - bo) <= 1e-10 * max(1.0, abs(bo))


# ---------------------------------------------------------------- generalised power cones (namespace cb::gp)
def _glib():
    L = _lib()
    i64p = C.POINTER(C.c_int64)
    L.gph_new.restype = C.c_void_p
    L.gph_new.argtypes = [C.c_int, C.POINTER(C.c_int32), i64p, i64p, f64p]
    L.gph_free.argtypes = [C.c_void_p]
    for nm in ["gph_m", "gph_nHs"]:
        getattr(L, nm).argtypes = [C.c_void_p]
    L.gph_unit_init.argtypes = [C.c_void_p, f64p, f64p]
    L.gph_update_scaling.argtypes = [C.c_void_p, f64p, C.c_double]
    L.gph_get_Hs.argtypes = [C.c_void_p, f64p, C.c_double]
    L.gph_mul_Hs.argtypes = [C.c_void_p, f64p, f64p]
    L.gph_copy_rows.argtypes = [C.c_void_p, f64p, f64p]
    L.gph_combined_shift.argtypes = [C.c_void_p, f64p, C.c_double]
    L.gph_step_length.argtypes = [C.c_void_p, f64p, f64p, f64p, f64p, C.c_double, C.c_double, C.c_double]
    L.gph_step_length.restype = C.c_double
    L.gph_barrier.argtypes = [C.c_void_p, f64p, f64p, f64p, f64p, C.c_double]
    L.gph_barrier.restype = C.c_double
    L.gph_kkt_values.argtypes = [C.c_void_p, f64p, f64p, f64p]
    return L


GP_CONES = [("genpow", ([0.6, 0.4], 1)), ("zero", 2), ("genpow", ([0.2, 0.3, 0.5], 2)), ("nonneg", 3),
            ("genpow", ([0.25, 0.25, 0.25, 0.25], 3)), ("genpow", ([1.0], 1))]


class GPair:
    def __init__(self, cones):
        self.L = _glib()
        codes = dict(CODES, genpow=6)
        ct = np.array([codes[k] for k, _ in cones], dtype=np.int32)
        cd = np.array([len(d[0]) if k == "genpow" else int(d) for k, d in cones], dtype=np.int64)
        c2 = np.array([int(d[1]) if k == "genpow" else 0 for k, d in cones], dtype=np.int64)
        al = np.array([a for k, d in cones if k == "genpow" for a in d[0]] or [0.0])
        self.h = self.L.gph_new(len(cones), ct.ctypes.data_as(C.POINTER(C.c_int32)), cd.ctypes.data_as(C.POINTER(C.c_int64)),
                                c2.ctypes.data_as(C.POINTER(C.c_int64)), P(al))
        self.m = self.L.gph_m(self.h)
        m = self.m
        self.ora = oracle.IPM(sp.csc_matrix((m, m)), np.zeros(m), -sp.identity(m, format="csc"), np.zeros(m), cones)
        self.rows = np.zeros(m, dtype=bool)
        self.blocks, self.offs = [], []
        o = 0
        for k, d in cones:
            rows = len(d[0]) + d[1] if k == "genpow" else d
            if k == "genpow":
                self.rows[o:o + rows] = True
                self.offs.append((o, len(d[0]), rows))
            o += rows

    def __del__(self):
        if getattr(self, "h", None):
            self.L.gph_free(self.h)

    def interior_point(self, rng, spread):
        z0, s0 = self.ora.unit_initialization()
        zero = np.zeros(self.m)
        for _ in range(200):
            z, s = z0 + spread * rng.standard_normal(self.m), s0 + spread * rng.standard_normal(self.m)
            if self.ora.update_scaling_ex(s, z, 1.0, 1) and np.isfinite(self.ora.compute_barrier(z, s, zero, zero, 0.0)):
                return z, s
        raise RuntimeError("no interior point found")


def test_genpow_unit_init_scaling_products_shift_and_kkt_columns():
    pr = GPair(GP_CONES)
    g = pr.rows
    z, s = np.zeros(pr.m), np.zeros(pr.m)
    pr.L.gph_unit_init(pr.h, P(z), P(s))
    zo, so = pr.ora.unit_initialization()
    assert np.array_equal(z[g], zo[g]) and np.array_equal(s[g], so[g])
    pr.ora.set_perm(np.arange(pr.ora.N))
    rng = np.random.default_rng(41)
    for trial in range(5):
        z, s = pr.interior_point(rng, 0.05 + 0.05 * trial)
        mu = 0.3 + 0.2 * trial
        assert pr.ora.update_scaling_ex(s, z, mu, 1)
        assert pr.L.gph_update_scaling(pr.h, P(z), mu) == 1
        Hs = np.zeros(pr.L.gph_nHs(pr.h))
        pr.L.gph_get_Hs(pr.h, P(Hs), -1.0)
        Ho = pr.ora.get_Hs()
        assert np.allclose(-Hs[g], Ho[g], rtol=1e-12, atol=0)      # all blocks here are diagonal: Hs index == row
        x = rng.standard_normal(pr.m)
        y = np.zeros(pr.m)
        pr.L.gph_mul_Hs(pr.h, P(y), P(x))
        assert np.allclose(y[g], pr.ora.mul_Hs(x)[g], rtol=1e-11, atol=1e-13) and np.all(y[~g] == 0.0)
        out = np.zeros(pr.m)
        pr.L.gph_copy_rows(pr.h, P(out), P(s))
        assert np.array_equal(out[g], pr.ora.affine_ds_ex(s)[g])
        sh = np.zeros(pr.m)
        pr.L.gph_combined_shift(pr.h, P(sh), 0.37 * mu)
        assert np.allclose(sh[g], pr.ora.combined_ds_shift(x, x, 0.37 * mu)[g], rtol=1e-12, atol=0)
        # the three expansion columns and their diagonals as KKTSolver::update writes them
        assert pr.ora.kkt_update()
        nz = pr.ora.kkt()[3]
        qr, pp, D = np.zeros(pr.m), np.zeros(pr.m), np.zeros(3 * len(pr.offs))
        pr.L.gph_kkt_values(pr.h, P(qr), P(pp), P(D))
        for k, (o, d1, rows) in enumerate(pr.offs):
            assert np.allclose(qr[o:o + d1], nz[pr.ora.genpow_map(k, "q")], rtol=1e-12, atol=0)
            assert np.allclose(qr[o + d1:o + rows], nz[pr.ora.genpow_map(k, "r")], rtol=1e-12, atol=0)
            assert np.allclose(pp[o:o + rows], nz[pr.ora.genpow_map(k, "p")], rtol=1e-12, atol=0)
            assert np.array_equal(D[3 * k:3 * k + 3], nz[pr.ora.genpow_map(k, "D")]) and list(D[3 * k:3 * k + 3]) == [-1.0, -1.0, 1.0]


def test_genpow_step_length_and_barrier():
    pr = GPair(GP_CONES)
    g = pr.rows
    rng = np.random.default_rng(43)
    seen = set()
    for trial in range(30):
        z, s = pr.interior_point(rng, 0.1)
        assert pr.L.gph_update_scaling(pr.h, P(z), 0.7) == 1 and pr.ora.update_scaling_ex(s, z, 0.7, 1)
        scale = [0.3, 1.0, 3.0, 10.0][trial % 4]
        dz, ds = scale * rng.standard_normal(pr.m), scale * rng.standard_normal(pr.m)
        dz[~g] = 0.0; ds[~g] = 0.0
        for amax in (1.0, 0.61):
            a = pr.L.gph_step_length(pr.h, P(dz), P(ds), P(z), P(s), amax, 1e-4, 0.8)
            assert a == pr.ora.step_length(dz, ds, z, s, amax)
            seen.add(a)
        dz *= 0.02 / scale; ds *= 0.02 / scale
        for al in (0.0, 0.5):
            b = pr.L.gph_barrier(pr.h, P(z), P(s), P(dz), P(ds), al)
            bo = pr.ora.compute_barrier(z, s, dz, ds, al)
            # the oracle sums every cone: take the nonnegative rows out again (rows 10..12 of GP_CONES: 3 + 2 + 5 = 10;
            # the zero cone has no barrier)
            nn = slice(10, 13)
            bo -= -np.sum(np.log((s[nn] + al * ds[nn]) * (z[nn] + al * dz[nn])))
            assert np.isfinite(bo) and abs(b - bo) <= 1e-10 * max(1.0, abs(bo))
    assert len(seen) > 3
