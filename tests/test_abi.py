"""CPU: the C-ABI library loads and exports every symbol include/clarabel_b200.h
declares; constructing a device object without a GPU fails loudly (no fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import clarabel_rs_b200 as cb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "clarabel_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(c(?:ldl|kkt|ipm|cone)_[A-Za-z0-9_]+)\s*\(", src)))


def test_all_declared_symbols_exported():
    L = cb.lib()
    names = declared_symbols()
    assert len(names) >= 15
    for nm in names:
        assert hasattr(L, nm), f"{nm} declared in the header but not exported"
    for nm in cb.EXPORTED_SYMBOLS:
        assert nm in names


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(cb.BackendError) as e:
        cb.CudaLDLSolver(2, [0, 1, 2], [0, 1], [1.0, -1.0], [1, -1])
    assert "Cuda" in str(e.value)


def test_default_opts_match_reference_settings():
    o = cb.cldl_opts()
    cb.lib().cldl_default_opts(C.byref(o))
    assert o.regularize_eps == 1e-13 and o.regularize_delta == 2e-7   # default/settings.rs:155-161
    assert o.regularize_enable == 1 and o.amd_dense_scale == 1.5       # ldlsolvers/qdldl.rs:38-41


def test_struct_mirrors_have_the_compiled_sizes():
    """the ctypes mirrors of the four structs that cross the ABI are as large as the C structs the library was built with
    (a mirror that is too small lets cldl_default_opts write past it)"""
    import ctypes as C
    import numpy as np
    L = cb.lib()
    out = np.zeros(4, dtype=np.uint64)
    L.cipm_abi_sizes.argtypes = [C.POINTER(C.c_uint64)]
    L.cipm_abi_sizes.restype = None
    L.cipm_abi_sizes(out.ctypes.data_as(C.POINTER(C.c_uint64)))
    assert [int(v) for v in out] == [C.sizeof(cb.pkg.cldl_opts), C.sizeof(cb.pkg.cldl_info_t), C.sizeof(cb.pkg.cipm_settings),
                                      C.sizeof(cb.pkg.cipm_info)]


def test_settable_infinity_bound():      # presolve.rs:107-114 (host state of the library: no device needed)
    cb.default_infinity()
    d = cb.get_infinity()
    cb.set_infinity(1e21)
    assert cb.get_infinity() == 1e21
    cb.default_infinity()
    assert cb.get_infinity() == d == 1e20
