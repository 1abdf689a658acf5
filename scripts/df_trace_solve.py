"""Task timeline of the dataflow triangular solves (k_solve_df) on a C2-like KKT.
Usage (GPU box): python scripts/df_trace_solve.py 100000 200000 200"""
import os, sys, struct
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("CB_DF_TRACE_SOLVE", "/tmp/df_trace_solve.bin")
import clarabel_rs_b200 as cb
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from helpers import workloads

n, m = int(sys.argv[1]), int(sys.argv[2])
window = None if sys.argv[3] == "None" else int(sys.argv[3])
pr = workloads.random_sparse_qp(n=n, m=m, nnz_per_row=5, seed=1, window=window)
N, cp, rv, nz, ds = workloads.kkt_triu(pr["P"], pr["A"], np.random.default_rng(0).uniform(0.5, 2, m))
s = cb.CudaLDLSolver(N, cp, rv, nz, ds, ordering=cb.ORDER_ND)
assert s.refactor()
b = np.random.default_rng(1).standard_normal(N)
for _ in range(3):
    x = s.solve(b)
print("solve_ms", s.time_solve_ms(10))
x = s.solve(b)
raw = open(os.environ["CB_DF_TRACE_SOLVE"], "rb").read()
nt = struct.unpack("q", raw[:8])[0]
front = np.frombuffer(raw, dtype=np.int32, count=nt, offset=8)
kind = np.frombuffer(raw, dtype=np.int32, count=nt, offset=8 + 4 * nt)
tr = np.frombuffer(raw, dtype=np.uint64, count=12 * nt, offset=8 + 8 * nt).reshape(2, nt, 6).astype(np.int64)
S = cb.SymbolicAnalysis(N, cp, rv, perm=s.perm())
nsv, nrv = np.diff(S.sn_first), np.diff(S.sn_rowptr)
for di, name in ((0, "forward"), (1, "backward")):
    t = tr[di]
    t0 = t[:, 0].min()
    grab, ready, end, m1, m2 = [(t[:, i] - t0) / 1e3 for i in (0, 1, 2, 3, 4)]
    print(f"== {name}: span {end.max():.1f} us, tasks {nt}")
    for k, kn in ((0, "narrow x8"), (1, "wide")):
        q = kind == k
        if not q.any():
            continue
        rd = np.where(t[:, 1][q] > 0, ready[q], grab[q])
        print(f"  {kn:9s} n={q.sum():6d} busy sum {np.sum(end[q]-rd)/1e3:8.2f} ms mean {np.mean(end[q]-rd):6.2f} us | wait mean {np.mean(rd-grab[q]):6.2f} us")
        if k == 1:
            a, b_ = ("gather", "trsv") if di == 0 else ("stage x", "gemv")
            c = "gemv rows" if di == 0 else "trsv"
            print(f"     phases: {a} {np.mean(m1[q]-rd):5.2f}  {b_} {np.mean(m2[q]-m1[q]):5.2f}  {c} {np.mean(end[q]-m2[q]):5.2f} us")
            if di == 0:
                m5 = (t[:, 5] - t0) / 1e3
                print(f"     trsv split: load+loop {np.mean(m5[q]-m1[q]):5.2f}  stores+sync {np.mean(m2[q]-m5[q]):5.2f}")
    nb = 20
    edges = np.linspace(0, end.max(), nb + 1)
    rd = np.where(t[:, 1] > 0, ready, grab)
    print("  slice(us) busyCTAs waitingCTAs")
    for i in range(nb):
        lo, hi = edges[i], edges[i + 1]
        bz = np.clip(np.minimum(end, hi) - np.maximum(rd, lo), 0, None).sum() / (hi - lo)
        wt = np.clip(np.minimum(rd, hi) - np.maximum(grab, lo), 0, None).sum() / (hi - lo)
        print(f"  {lo:8.0f} {bz:8.1f} {wt:8.1f}")
    # last 25 wide tasks to finish: the chain at the top of the tree
    order = np.argsort(end)[-25:] if di == 0 else np.argsort(end)[:25]
    print("  front ns nr | ready end busy (m1-ready, m2-m1, end-m2)")
    for i in order:
        f = front[i]
        print(f"  {f:7d} ns={nsv[f]:3d} nr={nrv[f]:5d} kind={kind[i]} | {rd[i]:8.1f} {end[i]:8.1f} {end[i]-rd[i]:6.1f} ({m1[i]-rd[i]:5.1f} {m2[i]-m1[i]:5.1f} {end[i]-m2[i]:5.1f})")
