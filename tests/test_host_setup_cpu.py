"""Host side of cipm_create on the CPU: cone collapsing, inf-bound presolve, KKT assembly (pattern, signs, sparse
expansion columns), ordering, symbolic analysis and plan construction of the PRODUCT library run in a subprocess with
tests/host_harness/libfake_cudart.so preloaded (device memory = host memory, kernel launches dropped -- see that
file's header).  No numbers are checked here (no kernel runs); the KKT structure must equal the oracle's.
The GPU parity tests (-m gpu) check the same constructors with real kernels."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "tests", "host_harness", "libfake_cudart.so")

CHILD = r'''
import json, sys
import numpy as np, scipy.sparse as sp
sys.path.insert(0, ROOT); sys.path.insert(0, ROOT + "/tests")
import clarabel_rs_b200 as cb, oracle
import ref_problems as rp, test_oracle_nonsym as ns
from helpers import workloads

def structure(name, P, q, A, b, cones, **kw):
    dev = cb.CudaSolver(P, q, A, b, cones, settings=cb.default_settings(**kw) if kw else None)
    ora = oracle.IPM(P, q, A, b, cones, settings=oracle.default_settings(**kw) if kw else None)
    N, cp, rv, _, ds = dev.kkt()
    No, cpo, rvo, _, dso = ora.kkt()
    perm = dev.kkt_perm()
    li = dev.linear_solver_info()
    out = dict(name=name, N=int(N), No=int(No), same=bool(N == No and np.array_equal(cp, cpo) and np.array_equal(rv, rvo) and np.array_equal(ds, dso)),
               perm_ok=bool(np.array_equal(np.sort(perm), np.arange(N))), m_reduced=int(dev.m_reduced), m_reduced_o=int(ora.m_reduced),
               symmetric=bool(dev.cone_is_symmetric()), nnzL=int(li.nnzL))
    dev.close()
    return out

res = []
res.append(structure("qp", *rp.basic_qp()))
res.append(structure("socp", *rp.basic_socp()))
res.append(structure("exp", *ns.expcone_data()))
res.append(structure("mixed", *ns.mixed_conic_data()))
res.append(structure("genpow", *ns.genpow_data()))
pr = workloads.entropy_power_mix(40, 20, n_eq=3, seed=6)
res.append(structure("entropy_power_mix", pr["P"], pr["q"], pr["A"], pr["b"], pr["cones"]))
pr = workloads.portfolio_socp(n_assets=120, n_soc=6, soc_dim=9, block=30, seed=7)
res.append(structure("portfolio", pr["P"], pr["q"], pr["A"], pr["b"], pr["cones"]))
pr = workloads.block_sdp(n=60, n_psd=4, psd_dim=4, nnz_per_row=3, window=20, n_nonneg=10, seed=4)
res.append(structure("sdp", pr["P"], pr["q"], pr["A"], pr["b"], pr["cones"]))
# a bigger mix with generalised power cones of several shapes next to sparse SOCs
cones = [("nonneg", 5), ("genpow", ([0.2, 0.3, 0.5], 2)), ("soc", 9), ("genpow", ([0.5, 0.5], 1)), ("exp", 3), ("soc", 3), ("genpow", ([1.0], 4)), ("pow", 0.3)]
m = 5 + 5 + 9 + 3 + 3 + 3 + 5 + 3
rng = np.random.default_rng(0)
A = sp.random(m, 12, density=0.3, random_state=3, format="csc") + sp.vstack([sp.identity(12), sp.csc_matrix((m - 12, 12))]).tocsc()
res.append(structure("genpow_mix", sp.identity(12, format="csc"), rng.standard_normal(12), A, rng.standard_normal(m), cones))
# inf-bound presolve (tests/presolve.rs)
n = 3
P = sp.identity(n, format="csc"); A = (2.0 * sp.vstack([sp.identity(n), -sp.identity(n)])).tocsc()
c = np.array([3., -2., 1.])
for tag, idx, cn in [("presolve1", [3], [("nonneg", 3), ("nonneg", 3)]), ("presolve2", [4], [("zero", 2), ("nonneg", 4)]),
                     ("presolve3", [0, 1, 2], [("nonneg", 3), ("nonneg", 3)]), ("presolve_all", list(range(6)), [("nonneg", 3), ("nonneg", 3)])]:
    b = np.ones(2 * n); b[idx] = 1e30
    res.append(structure(tag, P, c, A, b, cn))
# a caller-supplied KKT permutation has the length of the assembled system (presolved rows gone, expansion columns in)
P6, c6, A6, b6, cones6 = ns.genpow_data()
b6 = np.concatenate([b6, [1e30, 2.0]]); A6 = sp.vstack([A6, sp.csc_matrix(np.array([[1., 0, 0, 0, 0, 0], [0, 1., 0, 0, 0, 0]]))]).tocsc()
cones6 = cones6 + [("nonneg", 2)]
Nexp = 6 + (8 + 1) + 6
devp = cb.CudaSolver(P6, c6, A6, b6, cones6, kkt_perm=np.arange(Nexp)[::-1].copy())
res.append(dict(name="user_perm", N=int(devp.N), No=Nexp, same=bool(devp.N == Nexp), perm_ok=bool(np.array_equal(np.sort(devp.kkt_perm()), np.arange(Nexp))),
                m_reduced=int(devp.m_reduced), m_reduced_o=9, symmetric=False, nnzL=0))
print("RESULT " + json.dumps(res))
'''


@pytest.fixture(scope="module")
def results():
    if not os.path.exists(SHIM):
        pytest.fail("tests/host_harness/libfake_cudart.so missing: run `make`")
    env = dict(os.environ, LD_PRELOAD=SHIM)
    out = subprocess.run([sys.executable, "-c", "ROOT = %r\n" % ROOT + CHILD], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-3000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("RESULT ")][-1]
    return {r["name"]: r for r in json.loads(line[7:])}


@pytest.mark.parametrize("name", ["qp", "socp", "exp", "mixed", "genpow", "entropy_power_mix", "portfolio", "sdp",
                                  "genpow_mix", "presolve1", "presolve2", "presolve3", "presolve_all", "user_perm"])
def test_kkt_structure_of_the_product_constructor_equals_the_oracle(results, name):
    r = results[name]
    assert r["same"], r
    assert r["perm_ok"] and r["m_reduced"] == r["m_reduced_o"]


def test_constructor_facts(results):
    assert results["qp"]["symmetric"] and results["portfolio"]["symmetric"] and results["sdp"]["symmetric"]
    assert not results["exp"]["symmetric"] and not results["genpow"]["symmetric"] and not results["mixed"]["symmetric"]
    assert results["genpow"]["N"] == 6 + 8 + 6                      # 3 expansion columns per generalised power cone
    assert results["presolve1"]["m_reduced"] == 5 and results["presolve3"]["m_reduced"] == 3 and results["presolve_all"]["m_reduced"] == 0


def test_cpp_header_mirror_compiles_and_drives_the_c_abi(tmp_path):
    """include/clarabel_b200.hpp (the reference's trait / solver names over the C ABI): builds with g++, constructs and
    'solves' under the CUDA-runtime stand-in (host plumbing only), and fails loudly without a device."""
    exe = str(tmp_path / "hpp_example")
    lib_dir = os.path.join(ROOT, "clarabel.rs_b200")
    cc = subprocess.run(["g++", "-std=c++17", "-Wall", "-Werror", "-O1", "-o", exe,
                         os.path.join(ROOT, "tests", "host_harness", "hpp_example.cpp"),
                         "-L" + lib_dir, "-lclarabel_b200", "-Wl,-rpath," + lib_dir], capture_output=True, text=True)
    assert cc.returncode == 0, cc.stderr[-2000:]
    run = subprocess.run([exe], env=dict(os.environ, LD_PRELOAD=SHIM), capture_output=True, text=True, timeout=120)
    assert run.returncode == 0 and "created: kkt nnzA=17" in run.stdout and "mixed cones: ok" in run.stdout, run.stdout + run.stderr
    import shutil
    if shutil.which("nvidia-smi") is None:          # no device: the product path must refuse, not fall back
        bare = subprocess.run([exe], capture_output=True, text=True, timeout=120)
        assert bare.returncode == 2 and "SolverError" in bare.stdout
