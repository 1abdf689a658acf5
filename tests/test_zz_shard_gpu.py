"""GPU: ONE factorisation split over several ranks (subtree sharding) against the oracle and against the
unsharded device path.  All ranks live in this process on one device (`ShardedLDLGroup`): the phases, the packed
exchanges and the per-rank task queues are exactly those of a multi-GPU run, only the transport is a device copy
instead of NCCL, so the single-GPU test box checks the whole sharded path.  The same module runs on the CUDA-on-CPU
emulated build in the CPU suite (tests/test_emu_cpu.py)."""
import numpy as np
import pytest

import clarabel_rs_b200 as cb
from helpers import small_kkt
from oracle import QDLDL

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("nranks", [2, 3, 4])
@pytest.mark.parametrize("n,m,window,leaf", [(600, 1000, 30, 40), (1500, 2500, 60, 100)])
def test_sharded_factorisation_matches_oracle_and_unsharded_path(n, m, window, leaf, nranks):
    N, cp, rv, nz, ds = small_kkt(n, m, seed=1, window=window, k=3)
    g = cb.ShardedLDLGroup(N, cp, rv, nz, ds, nranks, ordering=cb.ORDER_ND, nd_leaf=leaf)
    one = cb.CudaLDLSolver(N, cp, rv, nz, ds, ordering=cb.ORDER_ND, nd_leaf=leaf)
    assert g.refactor() and one.refactor()
    assert np.array_equal(g.perm(), one.perm())
    f = QDLDL((N, N), cp, rv, nz, g.perm(), dsigns=ds, regularize_eps=1e-13, regularize_delta=2e-7)
    rng = np.random.default_rng(0)
    for _ in range(2):
        b = rng.standard_normal(N)
        xs, x1, xo = g.solve(b), one.solve(b), f.solve(b)
        for x in xs:
            assert np.max(np.abs(x - xo)) <= 1e-9 * max(1.0, np.max(np.abs(xo)))
            # same kernels, same summation orders, front by front: the sharded solution is the unsharded one bit for bit
            assert np.array_equal(x, x1)
    assert g.counts() == (f.regularize_count, f.positive_inertia)
    # something was actually split: every rank owns x entries and at least one cut root crosses ranks
    L = g._L
    assert all(int(L.cldl_shard_count(g.ranks[0]._h, 2, r)) > 0 for r in range(nranks))
    assert sum(int(L.cldl_shard_count(g.ranks[0]._h, 0, r)) for r in range(nranks)) > 0
    g.close(); one.close()


def test_sharded_refactor_after_value_update_and_regularised_pivots():
    # the recipe of test_ldl_gpu.py::test_dynamic_regularisation_counts, under nested dissection so that it shards
    N, cp, rv, nz, ds = small_kkt(300, 450, seed=5, window=10)
    nz = nz.copy()
    dgl = cp[1:] - 1
    nz[dgl[300:330]] = 0.0          # zero Hs entries
    nz[dgl[:5]] = -1.0              # wrong-signed P diagonal: forces dynamic regularisation (qdldl.rs:645-651)
    kw = dict(ordering=cb.ORDER_ND, nd_leaf=30)
    g = cb.ShardedLDLGroup(N, cp, rv, nz, ds, 3, **kw)
    one = cb.CudaLDLSolver(N, cp, rv, nz, ds, **kw)
    f = QDLDL((N, N), cp, rv, nz, g.perm(), dsigns=ds, regularize_eps=1e-13, regularize_delta=2e-7)
    assert g.refactor() == one.refactor() == True
    li = one.linear_solver_info()
    assert g.counts() == (li.regularize_count, li.positive_inertia) == (f.regularize_count, f.positive_inertia)
    assert f.regularize_count > 0
    b = np.random.default_rng(1).standard_normal(N)
    x1 = one.solve(b)
    assert all(np.array_equal(x, x1) for x in g.solve(b))
    # new values through the DirectLDLSolver trait on every rank, then refactor again
    idx = np.arange(0, len(nz), 7)
    vals = nz[idx] * 1.01
    for s in g.ranks + [one]:
        s.update_values(idx, vals)
    assert g.refactor() == one.refactor() == True
    li = one.linear_solver_info()
    assert g.counts() == (li.regularize_count, li.positive_inertia)
    x1 = one.solve(b)
    assert all(np.array_equal(x, x1) for x in g.solve(b))
    g.close(); one.close()


def test_unsharded_entry_points_refuse_on_a_sharded_handle():
    N, cp, rv, nz, ds = small_kkt(200, 300, seed=2, window=20, k=3)
    s = cb.CudaLDLSolver(N, cp, rv, nz, ds, shard_nranks=2, shard_rank=0)
    with pytest.raises(cb.BackendError):
        s.refactor()
    s.close()
