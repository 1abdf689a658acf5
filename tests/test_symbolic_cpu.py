"""CPU tests of the host symbolic analysis (ordering, supernodes, relative maps,
level schedule): a numpy emulation of the device schedule (tests/mf_emulator.py)
must reproduce the oracle's factors and solutions on the same permutation."""
import numpy as np
import pytest
import scipy.sparse as sp

import clarabel_rs_b200 as cb
from clarabel_rs_b200 import pkg  # noqa: F401
import mf_emulator as mf
from oracle import QDLDL

import importlib.util, os, sys
_wl = importlib.util.spec_from_file_location(
    "workloads", os.path.join(os.path.dirname(cb.pkg.__file__), "workloads.py"))
workloads = importlib.util.module_from_spec(_wl)
_wl.loader.exec_module(workloads)


def small_kkt(n, m, seed, window=None):
    pr = workloads.random_sparse_qp(n=n, m=m, nnz_per_row=3, seed=seed, window=window, p_offdiag=n)
    rng = np.random.default_rng(seed + 100)
    h = rng.uniform(0.5, 2.0, size=m)
    return workloads.kkt_triu(pr["P"], pr["A"], h)


def check_perm(p, n):
    assert sorted(p.tolist()) == list(range(n))


@pytest.mark.parametrize("ordering", [cb.ORDER_AMD, cb.ORDER_ND, cb.ORDER_BEST])
@pytest.mark.parametrize("n,m,window", [(30, 50, None), (200, 350, 20), (400, 300, None)])
@pytest.mark.parametrize("max_panel", [4, 64])
def test_emulated_schedule_matches_oracle(ordering, n, m, window, max_panel):
    N, cp, rv, nz, ds = small_kkt(n, m, seed=n + m, window=window)
    S = cb.SymbolicAnalysis(N, cp, rv, ordering=ordering, max_panel=max_panel, nd_leaf=16)
    check_perm(S.perm, N)
    assert np.all(S.iperm[S.perm] == np.arange(N))
    # postordered etree: parent index larger than child
    assert all(S.parent[j] == -1 or S.parent[j] > j for j in range(N))
    # oracle on the same permutation
    f = QDLDL((N, N), cp, rv, nz, S.perm, dsigns=ds, regularize_eps=1e-13, regularize_delta=2e-7)
    assert f.nnzL == S.nnzL                      # simplicial column counts agree with qdldl's etree pass
    assert np.all(np.diff(f.Lp) == S.colcount)
    assert np.all(np.where(f.etree < 0, -1, f.etree) == S.parent)
    Lpan, D, regc = mf.factor(S, nz, ds[S.perm])
    assert regc == f.regularize_count
    assert np.allclose(D, f.D, rtol=1e-10, atol=0)
    # every oracle L entry must appear in the supernodal panels with the same value
    Lp, Li, Lx = f.Lp, f.Li, f.Lx
    col2sn = np.searchsorted(S.sn_first, np.arange(N), side="right") - 1
    for j in range(N):
        s = col2sn[j]
        fcol, ns = S.sn_first[s], S.sn_first[s + 1] - S.sn_first[s]
        nr = S.sn_rowptr[s + 1] - S.sn_rowptr[s]
        ld = ns + nr
        rows = S.sn_rows[S.sn_rowptr[s]:S.sn_rowptr[s + 1]]
        col = Lpan[S.panel_off[s] + (j - fcol) * ld: S.panel_off[s] + (j - fcol + 1) * ld]
        for q in range(Lp[j], Lp[j + 1]):
            i = Li[q]
            li = i - fcol if i < fcol + ns else ns + np.searchsorted(rows, i)
            assert abs(col[li] - Lx[q]) <= 1e-9 * max(1.0, abs(Lx[q]))
    rng = np.random.default_rng(0)
    b = rng.standard_normal(N)
    x = mf.solve(S, Lpan, D, b)
    xo = f.solve(b)
    assert np.max(np.abs(x - xo)) <= 1e-9 * max(1.0, np.max(np.abs(xo)))


def test_given_perm_is_respected_up_to_postorder():
    N, cp, rv, nz, ds = small_kkt(40, 60, seed=3)
    perm = np.random.default_rng(1).permutation(N)
    S = cb.SymbolicAnalysis(N, cp, rv, perm=perm)
    check_perm(S.perm, N)
    assert S.ordering_used == 0
    f = QDLDL((N, N), cp, rv, nz, perm, dsigns=ds)
    assert f.nnzL == S.nnzL  # postordering never changes fill


def test_structural_errors():
    with pytest.raises(cb.BackendError):
        cb.SymbolicAnalysis(3, [0, 1, 1, 3], [0, 0, 2])         # empty column
    with pytest.raises(cb.BackendError):
        cb.SymbolicAnalysis(3, [0, 2, 3, 4], [0, 2, 1, 2])      # entry below the diagonal
    with pytest.raises(cb.BackendError):
        cb.SymbolicAnalysis(3, [0, 1, 2, 3], [0, 1, 2], perm=[0, 0, 1])


def test_amd_quality_on_arrow_and_grid():
    # arrowhead: AMD must put the hub last (fill = 0 beyond the arrow itself)
    n = 50
    cp = [0, 1] + [1 + 2 * k for k in range(1, n)]
    rv = [0] + sum(([0, k] for k in range(1, n)), [])
    p = cb.order(n, cp, rv, cb.ORDER_AMD)
    check_perm(p, n)
    S = cb.SymbolicAnalysis(n, cp, rv, perm=p)
    assert S.nnzL == n - 1
    # 2-D grid Laplacian 30x30: AMD / ND fill far below the natural (banded) ordering
    g = 30
    idx = lambda i, j: i * g + j
    ent = []
    for i in range(g):
        for j in range(g):
            ent.append((idx(i, j), idx(i, j)))
            if i + 1 < g: ent.append((idx(i, j), idx(i + 1, j)))
            if j + 1 < g: ent.append((idx(i, j), idx(i, j + 1)))
    ent.sort(key=lambda e: (e[1], e[0]))
    N = g * g
    cp = np.zeros(N + 1, dtype=np.int64)
    for r, c in ent: cp[c + 1] += 1
    cp = np.cumsum(cp)
    rv = np.array([r for r, c in ent])
    nat = cb.SymbolicAnalysis(N, cp, rv, perm=np.arange(N)).nnzL
    amd = cb.SymbolicAnalysis(N, cp, rv, ordering=cb.ORDER_AMD)
    nd = cb.SymbolicAnalysis(N, cp, rv, ordering=cb.ORDER_ND, nd_leaf=32)
    assert amd.nnzL < 0.6 * nat
    assert nd.nnzL < 0.8 * nat
    assert nd.nlevels <= amd.nlevels * 2 + 50


@pytest.mark.parametrize("ordering", [cb.ORDER_AMD, cb.ORDER_ND])
def test_update_arena_is_safe_without_level_barriers(ordering):
    """Two update matrices may share arena space only if one front is a proper ancestor of the other's PARENT
    (then the dependency graph itself orders the reuse after the last read): the numeric phase runs as a
    dataflow graph with no level barriers."""
    N, cp, rv, nz, ds = small_kkt(1500, 2500, seed=5, window=40)
    S = cb.SymbolicAnalysis(N, cp, rv, ordering=ordering, max_panel=16, nd_leaf=32)
    nr = np.diff(S.sn_rowptr)
    size = nr * nr
    live = np.where(size > 0)[0]
    order = live[np.argsort(S.upd_off[live], kind="stable")]
    assert S.upd_total >= int((S.upd_off[live] + size[live]).max())

    def is_proper_ancestor(a, s):      # a above s
        s = S.sn_parent[s]
        while s >= 0:
            if s == a:
                return True
            s = S.sn_parent[s]
        return False

    active = []                        # sweep over offsets
    overlaps = 0
    for s in order:
        lo, hi = S.upd_off[s], S.upd_off[s] + size[s]
        active = [t for t in active if S.upd_off[t] + size[t] > lo]
        for t in active:
            overlaps += 1
            pt, ps = S.sn_parent[t], S.sn_parent[s]
            ok = (pt >= 0 and is_proper_ancestor(s, pt)) or (ps >= 0 and is_proper_ancestor(t, ps))
            assert ok, (s, t)
        active.append(s)
    assert overlaps > 0                # the arena does recycle space


# ---- subtree-to-rank mapping for one factorisation on several GPUs (csrc/symbolic.h ShardPlan) ----
@pytest.mark.parametrize("nranks", [1, 2, 4, 8])
def test_shard_plan_invariants(nranks):
    pr = workloads.random_sparse_qp(n=6000, m=12000, nnz_per_row=4, seed=5, window=60)
    N, cp, rv = workloads.kkt_triu(pr["P"], pr["A"], np.ones(pr["A"].shape[0]))[:3]
    S = cb.SymbolicAnalysis(N, cp, rv, ordering=cb.ORDER_ND, nd_leaf=200)
    P = cb.shard_plan(S, nranks)
    owner, par = P["owner"], S.sn_parent
    assert owner.min() >= -1 and owner.max() < nranks
    ns = np.diff(S.sn_first).astype(float)
    nr = np.diff(S.sn_rowptr).astype(float)
    fl = ns ** 3 / 3 + ns ** 2 * nr + ns * nr ** 2
    for s in range(S.nsup):
        p = par[s]
        if p < 0:
            continue
        if owner[p] >= 0:
            assert owner[s] == owner[p]          # an owned front owns its whole subtree
        if owner[s] == -1:
            assert owner[p] == -1                # the replicated top part is closed upwards
    assert np.isclose(fl.sum(), P["total_flops"], rtol=1e-12) and np.isclose(P["total_flops"], S.flops_stored, rtol=1e-9)
    per_rank = np.array([fl[owner == g].sum() for g in range(nranks)])
    assert np.isclose(per_rank.max(), P["max_rank_flops"], rtol=1e-12)
    assert np.isclose(fl[owner == -1].sum(), P["top_flops"], rtol=1e-12, atol=1e-6)
    cut = [s for s in range(S.nsup) if owner[s] >= 0 and par[s] >= 0 and owner[par[s]] == -1]
    assert P["exchange_doubles"] == sum(int(nr[s]) ** 2 for s in cut) and P["exchange_vec"] == sum(int(nr[s]) for s in cut)
    assert 1.0 - 1e-12 <= P["model_speedup"] <= nranks + 1e-9
    if nranks == 1:
        assert (owner == 0).all() and P["top_flops"] == 0.0
    else:
        assert P["model_speedup"] > 1.3          # a nested-dissection tree does shard
        assert len(set(owner[owner >= 0])) == nranks


def test_shard_plan_of_a_chain_is_replicas_only():
    """an arrow / chain-shaped tree has no independent subtrees: the plan says speed-up 1"""
    n = 400
    rows = np.concatenate([np.arange(n), np.arange(n - 1)])          # tridiagonal: etree is a path
    cols = np.concatenate([np.arange(n), np.arange(1, n)])
    K = sp.csc_matrix((np.ones(rows.size), (rows, cols)), shape=(n, n))
    K.sort_indices()
    S = cb.SymbolicAnalysis(n, K.indptr, K.indices, perm=np.arange(n))
    P = cb.shard_plan(S, 4)
    assert P["model_speedup"] <= 1.0 + 1e-9


def test_amd_matches_reference_pin():
    """src/qdldl/test.rs:124-129 -- the one ordering the reference pins: AMD (dense scale 1.5) of its 4 x 4 test matrix
    is perm = [3, 0, 1, 2], iperm = [1, 2, 3, 0].  The `amd` crate is third-party and absent from /root/reference; the
    repo's own AMD (csrc/ordering.cpp) must reproduce it, tie-breaks and output convention included."""
    Ap, Ai = [0, 1, 3, 6, 8], [0, 0, 1, 0, 1, 2, 2, 3]          # test_matrix_4x4, test.rs:5-21
    perm = cb.order(4, Ap, Ai, cb.ORDER_AMD, 1.5)
    assert perm.tolist() == [3, 0, 1, 2]
    iperm = np.empty(4, dtype=np.int64); iperm[perm] = np.arange(4)
    assert iperm.tolist() == [1, 2, 3, 0]


@pytest.mark.parametrize("nblocks,link_blocks,rule", [(12, 12, "degree"), (8, 4, "locality")])
def test_hub_separator_finds_the_linking_rows_of_a_block_angular_problem(nblocks, link_blocks, rule):
    """Round 2, csrc/ordering.cpp hub_separator: a block-angular QP (banded blocks + a few linking rows that touch random
    blocks) is a small-world graph -- no BFS level is thin.  The nested dissection must find the linking rows as its
    top separator (they end up LAST in the permutation) and fall into the blocks below it.  Two proposal rules feed the
    connectivity step: above-typical degree (rows that touch 12 blocks) and, when no degree threshold works (rows that
    touch 4 blocks have FEWER neighbours than an ordinary row), locality -- no two neighbours of a connector are
    adjacent or share another neighbour.  Either way exactly the 120 linking rows must come out."""
    from helpers import workloads
    pr = workloads.block_angular_qp(n=48_000, nblocks=nblocks, nlink=120, link_blocks=link_blocks, window=48, seed=5)
    A = pr["A"]
    n, m = A.shape[1], A.shape[0]
    N, cp, rv, _, _ = workloads.kkt_triu(pr["P"], A, np.ones(m))
    nd = cb.SymbolicAnalysis(N, cp, rv, ordering=cb.ORDER_ND)
    amd = cb.SymbolicAnalysis(N, cp, rv, ordering=cb.ORDER_AMD)
    assert sorted(nd.perm.tolist()) == list(range(N))
    link_vertices = set(range(n + m - 120, n + m))            # the linking rows are the last 120 rows of A
    tail = set(nd.perm[-120:].tolist())
    assert tail == link_vertices, (rule, len(tail & link_vertices), "of 120 linking rows at the end of the ordering")
    assert nd.flops < 1.05 * amd.flops and nd.nlevels * 2 < amd.nlevels      # never more work than minimum degree, a much shorter tree
    best = cb.SymbolicAnalysis(N, cp, rv, ordering=cb.ORDER_BEST)
    assert best.ordering_used == cb.ORDER_ND
