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


def test_integration_md_rust_mirrors_match_the_header():
    """INTEGRATION.md's `#[repr(C)]` mirrors are what a maintainer would paste: field count, order, types and total size
    must be those of the C structs (round 1 shipped a CldlOpts that was two fields short of cldl_opts)."""
    import ctypes as C
    import re
    import clarabel_rs_b200 as cb
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    rust = {"f64": C.c_double, "i32": C.c_int32, "u32": C.c_uint32, "u64": C.c_uint64, "i64": C.c_int64}

    def mirror(name):
        body = re.search(r"pub struct %s \{(.*?)\n\}" % name, text, re.S).group(1)
        body = re.sub(r"//[^\n]*", "", body)
        fields = []
        for decl in body.split(","):
            decl = decl.strip()
            if not decl:
                continue
            nm, ty = [t.strip() for t in decl.split(":")]
            arr = re.match(r"\[u8; (\d+)\]", ty)
            fields.append((nm, C.c_char * int(arr.group(1)) if arr else rust[ty]))
        return type("M_" + name, (C.Structure,), {"_fields_": fields})

    for rname, cstruct in (("CldlOpts", cb.cldl_opts), ("CldlInfo", cb.cldl_info_t)):
        m = mirror(rname)
        assert len(m._fields_) == len(cstruct._fields_), rname
        assert C.sizeof(m) == C.sizeof(cstruct), (rname, C.sizeof(m), C.sizeof(cstruct))
        for (_, t1), (_, t2) in zip(m._fields_, cstruct._fields_):
            assert C.sizeof(t1) == C.sizeof(t2), rname
    sz = (C.c_uint64 * 4)()
    cb.lib().cipm_abi_sizes(sz)
    assert sz[0] == C.sizeof(cb.cldl_opts) and sz[1] == C.sizeof(cb.cldl_info_t)
    assert "cipm_abi_sizes" in text and "size_of::<CldlOpts>()" in text
