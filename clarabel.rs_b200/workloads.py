"""Deterministic synthetic conic problems (numpy / scipy, host side).

These are the BASELINE.json configs as generators.  The
reference ships no problem generators or benchmark inputs (SURVEY.md section
4), so the shapes follow BASELINE.json's (n, m, nnz) and cone lists.

Every generator returns a dict with scipy CSC ``P`` (upper triangle), ``A``,
vectors ``q``, ``b`` and ``cones`` as a list of (kind, dim) with kind in
{"zero", "nonneg", "soc", "psd"} -- the same vocabulary as the reference's
SupportedConeT (/root/reference/src/solver/core/cones/supportedcone.rs:17-52).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp


def _sym_psd_sparse(n, nnz_off, rng, window=None):
    """diag(U[0.1,1]) + nnz_off strictly-upper entries, made diagonally dominant."""
    rows = rng.integers(0, n, size=nnz_off)
    if window is None:
        cols = rng.integers(0, n, size=nnz_off)
    else:
        cols = np.clip(rows + rng.integers(1, window + 1, size=nnz_off), 0, n - 1)
    keep = rows != cols
    r, c = np.minimum(rows[keep], cols[keep]), np.maximum(rows[keep], cols[keep])
    v = 0.01 * rng.standard_normal(r.size)
    off = sp.coo_matrix((v, (r, c)), shape=(n, n)).tocsc()
    off.sum_duplicates()
    absrow = np.asarray(abs(off).sum(axis=1)).ravel() + np.asarray(abs(off).sum(axis=0)).ravel()
    d = rng.uniform(0.1, 1.0, size=n) + absrow
    P = (off + sp.diags(d)).tocsc()
    P.sort_indices()
    return sp.triu(P, format="csc")


def random_sparse_qp(n=100_000, m=200_000, nnz_per_row=5, seed=1, window=None, p_offdiag=None):
    """Config C2: min 1/2 x'Px + q'x s.t. Ax + s = b, s >= 0  (Nonneg cone only).

    ``window=None`` draws the column indices of every row of A uniformly from
    all n columns (SURVEY 8d wording).  ``window=w`` draws them from a sliding
    window of w columns centred on the row's position: same (n, m, nnz), but the
    KKT graph then has locality (like discretised / staged / block-structured
    QPs) instead of being an expander whose factor is essentially dense.
    """
    rng = np.random.default_rng(seed)
    if p_offdiag is None:
        p_offdiag = 2 * n
    k = nnz_per_row
    if window is None:
        cols = rng.integers(0, n, size=(m, k))
        # resample duplicates within a row (rare for n >> k)
        for _ in range(8):
            srt = np.sort(cols, axis=1)
            dup = (srt[:, 1:] == srt[:, :-1]).any(axis=1)
            if not dup.any():
                break
            cols[dup] = rng.integers(0, n, size=(int(dup.sum()), k))
    else:
        centre = (np.arange(m, dtype=np.int64) * n) // m
        lo = np.clip(centre - window // 2, 0, max(n - window, 0))
        w = min(window, n)
        # k distinct offsets per row inside the window
        offs = np.argsort(rng.random((m, w)), axis=1)[:, :k] if w <= 64 else None
        if offs is None:
            offs = rng.integers(0, w, size=(m, k))
            for _ in range(8):
                srt = np.sort(offs, axis=1)
                dup = (srt[:, 1:] == srt[:, :-1]).any(axis=1)
                if not dup.any():
                    break
                offs[dup] = rng.integers(0, w, size=(int(dup.sum()), k))
        cols = lo[:, None] + offs
    rows = np.repeat(np.arange(m), k)
    vals = rng.standard_normal(m * k)
    A = sp.coo_matrix((vals, (rows, cols.ravel())), shape=(m, n)).tocsc()
    A.sum_duplicates()
    A.sort_indices()
    P = _sym_psd_sparse(n, p_offdiag, rng, window=None if window is None else max(window // 2, 2))
    q = rng.standard_normal(n)
    x0 = rng.standard_normal(n)
    b = A @ x0 + rng.uniform(0.1, 1.0, size=m)
    return dict(P=P, q=q, A=A, b=b, cones=[("nonneg", m)], name=f"random_sparse_qp(n={n},m={m},k={k},window={window},seed={seed})")


def kkt_triu(P, A, hdiag):
    """Upper-triangular CSC KKT matrix [[P, A'],[., -diag(hdiag)]] with a
    structural diagonal everywhere, plus Dsigns.  Only for LDL-level tests with
    diagonal scaling blocks; the general assembly is in the product (csrc/kkt)."""
    n, m = P.shape[0], A.shape[0]
    Pu = sp.triu(P, format="coo")
    diag_missing = np.setdiff1d(np.arange(n), Pu.row[Pu.row == Pu.col])
    r = np.concatenate([Pu.row, diag_missing])
    c = np.concatenate([Pu.col, diag_missing])
    v = np.concatenate([Pu.data, np.zeros(diag_missing.size)])
    At = A.T.tocoo()
    r = np.concatenate([r, At.row, n + np.arange(m)])
    c = np.concatenate([c, n + At.col, n + np.arange(m)])
    v = np.concatenate([v, At.data, -np.asarray(hdiag, dtype=float)])
    # build CSC by hand so explicit zeros survive
    order = np.lexsort((r, c))
    r, c, v = r[order], c[order], v[order]
    N = n + m
    colptr = np.zeros(N + 1, dtype=np.int64)
    np.add.at(colptr, c + 1, 1)
    colptr = np.cumsum(colptr)
    dsigns = np.concatenate([np.ones(n, dtype=np.int8), -np.ones(m, dtype=np.int8)])
    return N, colptr, r.astype(np.int64), v.astype(np.float64), dsigns


def portfolio_socp(n_assets=5000, n_soc=200, soc_dim=26, block=100, seed=2):
    """Config C3: min 1/2 x'Px - mu'x  s.t. 1'x = 1, x >= 0, ||G_k x_{S_k}|| <= sigma_k.

    P = blockdiag of dense SPD blocks (F F'/block + 0.1 I); every SOC cone couples
    soc_dim-1 contiguous assets through a dense G_k.
    """
    rng = np.random.default_rng(seed)
    n = n_assets
    blocks = []
    for _ in range(n // block):
        F = rng.standard_normal((block, block))
        blocks.append(F @ F.T / block + 0.1 * np.eye(block))
    rem = n - (n // block) * block
    if rem:
        F = rng.standard_normal((rem, rem))
        blocks.append(F @ F.T / rem + 0.1 * np.eye(rem))
    P = sp.triu(sp.block_diag(blocks, format="csc"), format="csc")
    q = -rng.uniform(0.0, 0.1, size=n)
    rows, cols, vals, b = [], [], [], []
    # zero cone: sum x = 1
    rows += [0] * n; cols += list(range(n)); vals += [1.0] * n; b.append(1.0)
    r = 1
    # nonneg: -x + s = 0
    rows += list(range(r, r + n)); cols += list(range(n)); vals += [-1.0] * n; b += [0.0] * n
    r += n
    k = soc_dim - 1
    for c in range(n_soc):
        start = (c * (n - k)) // max(n_soc - 1, 1) if n_soc > 1 else 0
        S = np.arange(start, start + k)
        G = rng.standard_normal((k, k)) / 5.0
        b.append(1.0)            # sigma_k ; first SOC row has no x dependence
        r += 1
        for i in range(k):
            rows += [r + i] * k; cols += S.tolist(); vals += (-G[i]).tolist()
        b += [0.0] * k
        r += k
    A = sp.coo_matrix((vals, (rows, cols)), shape=(r, n)).tocsc()
    A.sort_indices()
    cones = [("zero", 1), ("nonneg", n)] + [("soc", soc_dim)] * n_soc
    return dict(P=P, q=q, A=A, b=np.array(b), cones=cones,
                name=f"portfolio_socp(assets={n},soc={n_soc}x{soc_dim},seed={seed})")


def block_angular_qp(n=1_000_000, nblocks=64, rows_per_var=1.5, nlink=2000, link_blocks=16, window=64, seed=3):
    """Config C4: block-angular sparse QP.  nblocks diagonal blocks; every ordinary
    row touches 5-6 variables inside one block (within a sliding window, so the
    blocks have locality), 1/6 of the rows are equalities; nlink linking rows touch
    one variable in each of link_blocks random blocks."""
    rng = np.random.default_rng(seed)
    m = int(rows_per_var * n)
    bs = n // nblocks
    m_ord = m - nlink
    blk = (np.arange(m_ord, dtype=np.int64) * nblocks) // m_ord
    pos_in_blk = np.arange(m_ord, dtype=np.int64) - (blk * m_ord) // nblocks
    rows_in_blk = np.maximum(((blk + 1) * m_ord) // nblocks - (blk * m_ord) // nblocks, 1)
    centre = (pos_in_blk * bs) // rows_in_blk
    k = 6
    lo = np.clip(centre - window // 2, 0, max(bs - window, 0))
    offs = rng.integers(0, min(window, bs), size=(m_ord, k))
    cols = blk[:, None] * bs + lo[:, None] + offs
    keep = np.ones((m_ord, k), dtype=bool)
    keep[:, 5] = rng.random(m_ord) < 0.5        # 5 or 6 entries per row
    r = np.repeat(np.arange(m_ord), k)[keep.ravel()]
    cidx = cols.ravel()[keep.ravel()]
    v = rng.standard_normal(cidx.size)
    link_blocks = min(link_blocks, nblocks)
    lr = np.repeat(np.arange(m_ord, m), link_blocks)
    lb = np.stack([rng.choice(nblocks, size=link_blocks, replace=False) for _ in range(nlink)])
    lc = (lb * bs + rng.integers(0, bs, size=(nlink, link_blocks))).ravel()
    lv = rng.standard_normal(lc.size)
    A = sp.coo_matrix((np.concatenate([v, lv]), (np.concatenate([r, lr]), np.concatenate([cidx, lc]))), shape=(m, n)).tocsc()
    A.sum_duplicates(); A.sort_indices()
    # P: diagonal + one off-diagonal per column inside its block
    j = np.arange(n)
    jn = np.minimum(j + 1, ((j // bs) + 1) * bs - 1)
    off = sp.coo_matrix((0.05 * rng.standard_normal(n), (np.minimum(j, jn), np.maximum(j, jn))), shape=(n, n)).tocsc()
    off = sp.triu(off, k=1, format="csc")
    absr = np.asarray(abs(off).sum(axis=0)).ravel() + np.asarray(abs(off).sum(axis=1)).ravel()
    P = (off + sp.diags(rng.uniform(0.1, 1.0, n) + absr)).tocsc()
    P.sort_indices()
    q = rng.standard_normal(n)
    x0 = rng.standard_normal(n)
    nz = m // 6
    b = A @ x0
    b[nz:] += rng.uniform(0.1, 1.0, size=m - nz)
    cones = [("zero", nz), ("nonneg", m - nz)]
    return dict(P=sp.triu(P, format="csc"), q=q, A=A, b=b, cones=cones,
                name=f"block_angular_qp(n={n},blocks={nblocks},m={m},link={nlink},seed={seed})")


def block_sdp(n=20_000, n_psd=500, psd_dim=20, nnz_per_row=10, window=400, n_nonneg=2000, seed=4):
    """Config C5: block-diagonal SDP.  n_psd PSD(psd_dim) cones; every PSD row touches ~nnz_per_row variables of
    a window of `window` variables belonging to that cone; one trace-normalisation equality per cone; a few bound rows.
    min 0.005 x'x + q'x  s.t.  svec(F_k0) - A_k x in PSD,  a_k'x = b_k,  G x <= h."""
    rng = np.random.default_rng(seed)
    numel = psd_dim * (psd_dim + 1) // 2
    rows, cols, vals, b = [], [], [], []
    r = 0
    # zero cone rows
    for k in range(n_psd):
        lo = (k * max(n - window, 1)) // max(n_psd - 1, 1) if n_psd > 1 else 0
        cc = lo + rng.choice(min(window, n), size=min(nnz_per_row, n), replace=False)
        rows += [r] * cc.size; cols += cc.tolist(); vals += rng.standard_normal(cc.size).tolist()
        b.append(rng.standard_normal() * 0.1)
        r += 1
    nz = r
    # nonneg rows:  g'x + s = h, strictly feasible at x = 0
    for k in range(n_nonneg):
        lo = int(rng.integers(0, max(n - window, 1)))
        cc = lo + rng.choice(min(window, n), size=3, replace=False)   # local bound rows
        rows += [r] * 3; cols += cc.tolist(); vals += rng.standard_normal(3).tolist()
        b.append(rng.uniform(0.5, 1.5))
        r += 1
    # PSD rows:  s = svec(F0) - A x  with F0 = identity-ish (strictly feasible at x = 0)
    diag_idx = [k * (k + 3) // 2 for k in range(psd_dim)]
    for k in range(n_psd):
        lo = (k * max(n - window, 1)) // max(n_psd - 1, 1) if n_psd > 1 else 0
        for t in range(numel):
            cc = lo + rng.choice(min(window, n), size=min(nnz_per_row, n), replace=False)
            rows += [r + t] * cc.size; cols += cc.tolist(); vals += (0.3 * rng.standard_normal(cc.size)).tolist()
        bb = np.zeros(numel); bb[diag_idx] = 1.0
        b += bb.tolist()
        r += numel
    A = sp.coo_matrix((vals, (rows, cols)), shape=(r, n)).tocsc()
    A.sum_duplicates(); A.sort_indices()
    P = sp.identity(n, format="csc") * 0.01
    q = rng.standard_normal(n) * 0.1
    cones = [("zero", nz), ("nonneg", n_nonneg)] + [("psd", psd_dim)] * n_psd
    return dict(P=P, q=q, A=A, b=np.array(b), cones=cones,
                name=f"block_sdp(n={n},psd={n_psd}x{psd_dim},seed={seed})")


def entropy_power_mix(k_exp=200, k_pow=100, n_eq=10, seed=6):
    """Nonsymmetric-cone workload: entropy maximisation over a probability vector with linear moment constraints
    (k_exp exponential cones) next to a weighted geometric-mean allocation (k_pow 3-D power cones with exponents
    drawn from (0.1, 0.9)), sharing one budget row.

        max  sum_i t_i + sum_j c_j y_j
        s.t. (t_i, p_i, 1) in K_exp                  (t_i <= -p_i log p_i)
             sum p = 1,  F p = F p0                  (zero cone, p0 the uniform distribution perturbed)
             (u_j, w_j, y_j) in K_pow(alpha_j)       (|y_j| <= u_j^alpha_j w_j^(1-alpha_j))
             sum u + sum w + s = k_pow, s >= 0       (budget, nonnegative cone)

    cones use the vocabulary of the reference's SupportedConeT (supportedcone.rs:17-52): ("exp", 3), ("pow", alpha).
    """
    rng = np.random.default_rng(seed)
    n = 2 * k_exp + 3 * k_pow
    ip, it = 0, k_exp                       # p, t
    iu, iw, iy = 2 * k_exp, 2 * k_exp + k_pow, 2 * k_exp + 2 * k_pow
    rows, cols, vals, b, cones = [], [], [], [], []
    r = 0
    # equality rows
    p0 = rng.uniform(0.5, 1.5, k_exp); p0 /= p0.sum()
    rows += [r] * k_exp; cols += list(range(ip, ip + k_exp)); vals += [1.0] * k_exp; b.append(1.0); r += 1
    F = rng.standard_normal((n_eq, k_exp))
    for e in range(n_eq):
        rows += [r] * k_exp; cols += list(range(ip, ip + k_exp)); vals += F[e].tolist(); b.append(float(F[e] @ p0)); r += 1
    cones.append(("zero", 1 + n_eq))
    # budget row
    rows += [r] * (2 * k_pow); cols += list(range(iu, iu + 2 * k_pow)); vals += [1.0] * (2 * k_pow); b.append(float(k_pow)); r += 1
    cones.append(("nonneg", 1))
    # exponential cones: s = b - A x = (t_i, p_i, 1)
    for i in range(k_exp):
        rows += [r, r + 1]; cols += [it + i, ip + i]; vals += [-1.0, -1.0]; b += [0.0, 0.0, 1.0]; r += 3
        cones.append(("exp", 3))
    # power cones: s = (u_j, w_j, y_j)
    alphas = rng.uniform(0.1, 0.9, k_pow)
    for j in range(k_pow):
        rows += [r, r + 1, r + 2]; cols += [iu + j, iw + j, iy + j]; vals += [-1.0, -1.0, -1.0]; b += [0.0, 0.0, 0.0]; r += 3
        cones.append(("pow", float(alphas[j])))
    A = sp.coo_matrix((vals, (rows, cols)), shape=(r, n)).tocsc()
    A.sum_duplicates(); A.sort_indices()
    q = np.zeros(n)
    q[it:it + k_exp] = -1.0
    q[iy:iy + k_pow] = -rng.uniform(0.5, 1.5, k_pow)
    P = sp.csc_matrix((n, n))
    return dict(P=P, q=q, A=A, b=np.array(b), cones=cones,
                name=f"entropy_power_mix(exp={k_exp},pow={k_pow},seed={seed})")
