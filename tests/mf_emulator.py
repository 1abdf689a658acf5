"""numpy emulation of the device multifrontal schedule (same data structures,
same order of operations as clarabel.rs_b200/csrc/ldl.cu), used on CPU to
validate the host symbolic analysis before anything runs on a GPU."""
import numpy as np


def factor(S, vals, dsigns_perm, eps=1e-13, delta=2e-7, reg=True):
    n, nsup = S.n, S.nsup
    Lpan = np.zeros(S.L_alloc)
    U = np.zeros(max(S.upd_total, 1))
    D = np.zeros(n)
    regcount = 0
    for l in range(S.nlevels):
        for t in range(S.level_ptr[l], S.level_ptr[l + 1]):
            s = S.level_tasks[t]
            f, ns = S.sn_first[s], S.sn_first[s + 1] - S.sn_first[s]
            nr = S.sn_rowptr[s + 1] - S.sn_rowptr[s]
            ld = ns + nr
            W = np.zeros((ld, ns), order="F")
            Us = np.zeros((nr, nr), order="F")
            a0, a1 = S.asm_ptr[s], S.asm_ptr[s + 1]
            Wf = W.reshape(-1, order="F")
            Wf[S.asm_dst[a0:a1]] = vals[S.asm_src[a0:a1]]
            W = Wf.reshape((ld, ns), order="F")
            for c in S.child_list[S.child_ptr[s]:S.child_ptr[s + 1]]:
                nrc = S.sn_rowptr[c + 1] - S.sn_rowptr[c]
                Uc = U[S.upd_off[c]:S.upd_off[c] + nrc * nrc].reshape((nrc, nrc), order="F")
                rel = S.rel[S.sn_rowptr[c]:S.sn_rowptr[c + 1]]
                for b in range(nrc):
                    rb = rel[b]
                    for a in range(b, nrc):
                        ra = rel[a]
                        if rb < ns:
                            W[ra, rb] += Uc[a, b]
                        else:
                            Us[ra - ns, rb - ns] += Uc[a, b]
            for j in range(ns):
                dj = W[j, j]
                if reg:
                    sg = float(dsigns_perm[f + j])
                    if dj * sg < eps:
                        dj = delta * sg
                        regcount += 1
                D[f + j] = dj
                inv = 1.0 / dj
                for k in range(j + 1, ns):
                    wk = W[k, j] * inv
                    W[k:, k] -= W[k:, j] * wk
                W[j + 1:, j] *= inv
            L21 = W[ns:, :]
            Us -= np.tril((L21 * D[f:f + ns]) @ L21.T)
            Lpan[S.panel_off[s]:S.panel_off[s] + ld * ns] = W.reshape(-1, order="F")
            U[S.upd_off[s]:S.upd_off[s] + nr * nr] = Us.reshape(-1, order="F")
    return Lpan, D, regcount


def solve(S, Lpan, D, b):
    n = S.n
    xp = b[S.perm].astype(float).copy()
    u = np.zeros(max(len(S.sn_rows), 1))
    for l in range(S.nlevels):
        for t in range(S.level_ptr[l], S.level_ptr[l + 1]):
            s = S.level_tasks[t]
            f, ns = S.sn_first[s], S.sn_first[s + 1] - S.sn_first[s]
            rp = S.sn_rowptr[s]
            nr = S.sn_rowptr[s + 1] - rp
            ld = ns + nr
            P = Lpan[S.panel_off[s]:S.panel_off[s] + ld * ns].reshape((ld, ns), order="F")
            us = np.zeros(nr)
            for c in S.child_list[S.child_ptr[s]:S.child_ptr[s + 1]]:
                crp, nrc = S.sn_rowptr[c], S.sn_rowptr[c + 1] - S.sn_rowptr[c]
                rel = S.rel[crp:crp + nrc]
                for a in range(nrc):
                    if rel[a] < ns:
                        xp[f + rel[a]] += u[crp + a]
                    else:
                        us[rel[a] - ns] += u[crp + a]
            y = xp[f:f + ns].copy()
            for j in range(ns):
                y[j + 1:] -= P[j + 1:ns, j] * y[j]
            xp[f:f + ns] = y
            us -= P[ns:, :] @ y
            u[rp:rp + nr] = us
    out = np.zeros(n)
    for l in range(S.nlevels - 1, -1, -1):
        for t in range(S.level_ptr[l], S.level_ptr[l + 1]):
            s = S.level_tasks[t]
            f, ns = S.sn_first[s], S.sn_first[s + 1] - S.sn_first[s]
            rp = S.sn_rowptr[s]
            nr = S.sn_rowptr[s + 1] - rp
            ld = ns + nr
            P = Lpan[S.panel_off[s]:S.panel_off[s] + ld * ns].reshape((ld, ns), order="F")
            rows = S.sn_rows[rp:rp + nr]
            t_ = xp[f:f + ns] / D[f:f + ns] - P[ns:, :].T @ xp[rows]
            for j in range(ns - 1, 0, -1):
                t_[:j] -= P[j, :j] * t_[j]
            xp[f:f + ns] = t_
            out[S.perm[f:f + ns]] = t_
    return out
