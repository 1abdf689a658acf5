"""ORACLE -- TEST INFRASTRUCTURE ONLY (ctypes view of oracle/liboracle.so).

CPU restatement of the reference's hot path.  May be imported only by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None
i64p = C.POINTER(C.c_int64)
f64p = C.POINTER(C.c_double)
i8p = C.POINTER(C.c_int8)


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            raise RuntimeError("oracle/liboracle.so missing: run `make` (or __graft_entry__.build())")
        L = C.CDLL(path)
        vp = C.c_void_p
        L.oq_new.argtypes = [C.POINTER(vp), C.c_int64, C.c_int64, i64p, i64p, f64p, i64p, i8p,
                             C.c_int, C.c_int, C.c_double, C.c_double]
        L.oq_free.argtypes = [vp]
        L.oq_free.restype = None
        L.oq_refactor.argtypes = [vp]
        L.oq_solve.argtypes = [vp, f64p]
        L.oq_update_values.argtypes = [vp, i64p, f64p, C.c_int64]
        L.oq_update_values.restype = None
        L.oq_scale_values.argtypes = [vp, i64p, C.c_int64, C.c_double]
        L.oq_scale_values.restype = None
        L.oq_offset_values.argtypes = [vp, i64p, C.c_int64, C.c_double, i8p]
        L.oq_offset_values.restype = None
        L.oq_dinv_is_finite.argtypes = [vp]
        for nm in ["oq_n", "oq_nnzA", "oq_nnzL", "oq_regularize_count", "oq_positive_inertia"]:
            getattr(L, nm).argtypes = [vp]
            getattr(L, nm).restype = C.c_int64
        for nm in ["oq_Lp", "oq_Li", "oq_etree_ptr", "oq_Lnz", "oq_permA_colptr", "oq_permA_rowval",
                   "oq_AtoPAPt"]:
            getattr(L, nm).argtypes = [vp]
            getattr(L, nm).restype = i64p
        for nm in ["oq_Lx", "oq_D", "oq_Dinv", "oq_permA_nzval"]:
            getattr(L, nm).argtypes = [vp]
            getattr(L, nm).restype = f64p
        L.oq_invperm.argtypes = [C.c_int64, i64p, i64p]
        L.oq_permute.argtypes = [C.c_int64, f64p, f64p, i64p]
        L.oq_permute.restype = None
        L.oq_ipermute.argtypes = [C.c_int64, f64p, f64p, i64p]
        L.oq_ipermute.restype = None
        L.oq_permute_symmetric.argtypes = [C.c_int64, i64p, i64p, f64p, i64p, i64p, i64p, f64p, i64p]
        L.oq_permute_symmetric.restype = None
        L.oq_etree.argtypes = [C.c_int64, i64p, i64p, i64p, i64p, i64p]
        L.oq_etree.restype = None
        for nm in ["oq_lsolve", "oq_ltsolve"]:
            getattr(L, nm).argtypes = [C.c_int64, i64p, i64p, f64p, f64p]
            getattr(L, nm).restype = None
        L.oq_dltsolve.argtypes = [C.c_int64, i64p, i64p, f64p, f64p, f64p]
        L.oq_dltsolve.restype = None
        L.oq_solve_factors.argtypes = [C.c_int64, i64p, i64p, f64p, f64p, f64p]
        L.oq_solve_factors.restype = None
        _lib = L
    return _lib


def I(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def F(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def P(a):
    if a.dtype == np.int64:
        return a.ctypes.data_as(i64p)
    if a.dtype == np.float64:
        return a.ctypes.data_as(f64p)
    if a.dtype == np.int8:
        return a.ctypes.data_as(i8p)
    raise TypeError(a.dtype)


class QDLDLError(Exception):
    def __init__(self, code):
        names = {-1: "IncompatibleDimension", -2: "EmptyColumn", -3: "NotUpperTriangular",
                 -4: "ZeroPivot", -5: "InvalidPermutation"}
        super().__init__(names.get(code, str(code)))
        self.code = code


class QDLDL:
    """Oracle factorisation object (mirrors QDLDLFactorisation, qdldl.rs:72-211)."""

    def __init__(self, shape, colptr, rowval, nzval, perm, dsigns=None, logical=False,
                 regularize_enable=True, regularize_eps=1e-12, regularize_delta=1e-7):
        L = lib()
        self._L = L
        m, n = shape
        cp, rv, nz, pm = I(colptr), I(rowval), F(nzval), I(perm)
        ds = np.ascontiguousarray(dsigns, dtype=np.int8) if dsigns is not None else None
        h = C.c_void_p()
        rc = L.oq_new(C.byref(h), m, n, P(cp), P(rv), P(nz), P(pm), P(ds) if ds is not None else None,
                      1 if logical else 0, 1 if regularize_enable else 0, regularize_eps, regularize_delta)
        if rc:
            raise QDLDLError(rc)
        self._h = h
        self.n = n

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.oq_free(self._h)
            self._h = None

    def refactor(self):
        rc = self._L.oq_refactor(self._h)
        if rc:
            raise QDLDLError(rc)

    def refactor_ok(self):
        """Adapter-level bool (ldlsolvers/qdldl.rs:99-106)."""
        self.refactor()
        return bool(self._L.oq_dinv_is_finite(self._h))

    def solve(self, b):
        x = F(b).copy()
        if self._L.oq_solve(self._h, P(x)) != 0:
            raise AssertionError("solve on a logical factorisation (qdldl.rs:118)")
        return x

    def update_values(self, index, values):
        idx, v = I(index), F(values)
        self._L.oq_update_values(self._h, P(idx), P(v), idx.size)

    def scale_values(self, index, scale):
        idx = I(index)
        self._L.oq_scale_values(self._h, P(idx), idx.size, float(scale))

    def offset_values(self, index, offset, signs):
        idx = I(index)
        sg = np.ascontiguousarray(signs, dtype=np.int8)
        assert idx.size == sg.size
        self._L.oq_offset_values(self._h, P(idx), idx.size, float(offset), P(sg))

    def _arr(self, fn, n, dt):
        p = getattr(self._L, fn)(self._h)
        return np.ctypeslib.as_array(p, shape=(max(int(n), 1),))[:int(n)].astype(dt, copy=True)

    @property
    def nnzL(self):
        return int(self._L.oq_nnzL(self._h))

    @property
    def nnzA(self):
        return int(self._L.oq_nnzA(self._h))

    @property
    def regularize_count(self):
        return int(self._L.oq_regularize_count(self._h))

    @property
    def positive_inertia(self):
        return int(self._L.oq_positive_inertia(self._h))

    @property
    def D(self):
        return self._arr("oq_D", self.n, np.float64)

    @property
    def Dinv(self):
        return self._arr("oq_Dinv", self.n, np.float64)

    @property
    def Lp(self):
        return self._arr("oq_Lp", self.n + 1, np.int64)

    @property
    def Li(self):
        return self._arr("oq_Li", self.nnzL, np.int64)

    @property
    def Lx(self):
        return self._arr("oq_Lx", self.nnzL, np.float64)

    @property
    def etree(self):
        return self._arr("oq_etree_ptr", self.n, np.int64)

    @property
    def permA(self):
        cp = self._arr("oq_permA_colptr", self.n + 1, np.int64)
        return (cp, self._arr("oq_permA_rowval", self.nnzA, np.int64),
                self._arr("oq_permA_nzval", self.nnzA, np.float64))

    @property
    def AtoPAPt(self):
        return self._arr("oq_AtoPAPt", self.nnzA, np.int64)
