/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY.  Not product code.
 *
 * CPU restatement (plain C, single thread, f64, 64-bit indices) of the
 * reference's quasidefinite LDL^T path, /root/reference/src/qdldl/qdldl.rs.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library; the product path
 * (clarabel.rs_b200/) never links or calls it.
 *
 * Parity pinning: checked in tests/test_oracle_qdldl.py against every
 * known-answer vector of the reference's own src/qdldl/test.rs (exact
 * equality where the reference asserts exact equality) and against the
 * trait-level golden of ldlsolvers/faer_ldl.rs:352-409.
 *
 * The AMD ordering lives in an un-vendored third-party crate (amd 0.2.2,
 * Cargo.toml:18) and is NOT restated here: the oracle always takes an
 * explicit permutation (the reference supports this, qdldl.rs:36-38,243-246).
 *
 * Function -> reference map
 *   oq_invperm            qdldl.rs:771-782   (_invperm)
 *   oq_permute/ipermute   qdldl.rs:789-801
 *   oq_permute_symmetric  qdldl.rs:806-903   (Davis 2-pass, unsorted columns)
 *   oq_etree              qdldl.rs:433-464
 *   oq_factor             qdldl.rs:469-669   (_factor_inner, up-looking)
 *   oq_lsolve/oq_ltsolve/oq_dltsolve/oq_solve_factors   qdldl.rs:708-768
 *   oq_new/oq_refactor/oq_solve/oq_update_values/
 *   oq_scale_values/oq_offset_values                    qdldl.rs:95-211,230-295
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

typedef int64_t idx;
#define OQ_UNKNOWN ((idx)-1) /* stands for usize::MAX, qdldl.rs:426 */

/* error codes (mirror QDLDLError, qdldl.rs:10-26) */
enum { OQ_OK = 0, OQ_ERR_DIM = -1, OQ_ERR_EMPTYCOL = -2, OQ_ERR_NOT_TRIU = -3,
       OQ_ERR_ZERO_PIVOT = -4, OQ_ERR_BAD_PERM = -5 };

typedef struct {
    idx n;
    idx *perm, *iperm;
    /* permuted triu copy of the input + entry map (workspace.triuA / AtoPAPt) */
    idx *Ap, *Ai; double *Ax; idx nnzA; idx *AtoPAPt;
    /* factors */
    idx *Lp, *Li; double *Lx; idx nnzL;
    double *D, *Dinv;
    /* workspace */
    idx *etree, *Lnz, *iwork; unsigned char *bwork; double *fwork;
    int8_t *Dsigns;
    int regularize_enable; double regularize_eps, regularize_delta;
    idx regularize_count, positive_inertia;
    int is_symbolic;
} oq_t;

/* ---- permutation helpers ------------------------------------------------ */

int oq_invperm(idx n, const idx *p, idx *out)
{
    /* mirrors the reference's test `b[j]==0` uniqueness check, including its
       blind spot for index 0 (qdldl.rs:774-779) */
    for (idx i = 0; i < n; i++) out[i] = 0;
    for (idx i = 0; i < n; i++) {
        idx j = p[i];
        if (j >= 0 && j < n && out[j] == 0) out[j] = i;
        else return OQ_ERR_BAD_PERM;
    }
    return OQ_OK;
}

void oq_permute(idx n, double *x, const double *b, const idx *p)
{ for (idx i = 0; i < n; i++) x[i] = b[p[i]]; }

void oq_ipermute(idx n, double *x, const double *b, const idx *p)
{ for (idx i = 0; i < n; i++) x[p[i]] = b[i]; }

/* Symmetric permutation of a triu CSC matrix to another triu CSC matrix.
   Two passes: count entries per destination column, then fill in the order
   the source is traversed (so destination columns are NOT row-sorted). */
void oq_permute_symmetric(idx n, const idx *Ac, const idx *Ar, const double *Av,
                          const idx *iperm, idx *Pc, idx *Pr, double *Pv,
                          idx *AtoPAPt)
{
    idx *cnt = (idx *)calloc((size_t)(n > 0 ? n : 1), sizeof(idx));
    for (idx ca = 0; ca < n; ca++) {
        idx cp = iperm[ca];
        for (idx k = Ac[ca]; k < Ac[ca + 1]; k++) {
            idx ra = Ar[k];
            if (ra <= ca) {
                idx rp = iperm[ra];
                cnt[rp > cp ? rp : cp] += 1;
            }
        }
    }
    Pc[0] = 0;
    for (idx j = 0; j < n; j++) Pc[j + 1] = Pc[j] + cnt[j];
    for (idx j = 0; j < n; j++) cnt[j] = Pc[j]; /* next free slot per column */
    for (idx ca = 0; ca < n; ca++) {
        idx cp = iperm[ca];
        for (idx k = Ac[ca]; k < Ac[ca + 1]; k++) {
            idx ra = Ar[k];
            if (ra <= ca) {
                idx rp = iperm[ra];
                idx col = rp > cp ? rp : cp;
                idx dst = cnt[col]++;
                Pr[dst] = rp < cp ? rp : cp;
                Pv[dst] = Av[k];
                AtoPAPt[k] = dst;
            }
        }
    }
    free(cnt);
}

/* ---- symbolic ------------------------------------------------------------ */

void oq_etree(idx n, const idx *Ap, const idx *Ai, idx *work, idx *Lnz, idx *etree)
{
    for (idx i = 0; i < n; i++) { work[i] = 0; Lnz[i] = 0; etree[i] = OQ_UNKNOWN; }
    for (idx j = 0; j < n; j++) {
        work[j] = j;
        for (idx p = Ap[j]; p < Ap[j + 1]; p++) {
            idx i = Ai[p];
            while (work[i] != j) {
                if (etree[i] == OQ_UNKNOWN) etree[i] = j;
                Lnz[i] += 1;
                work[i] = j;
                i = etree[i];
            }
        }
    }
}

/* ---- numeric up-looking factorisation ----------------------------------- */

int oq_factor(idx n, const idx *Ap, const idx *Ai, const double *Ax,
              idx *Lp, idx *Li, double *Lx, double *D, double *Dinv,
              const idx *Lnz, const idx *etree,
              unsigned char *ymark, idx *iwork, double *yval,
              int logical, const int8_t *Dsigns,
              int reg_enable, double reg_eps, double reg_delta,
              idx *reg_count, idx *pos_inertia)
{
    idx *yidx = iwork, *ebuf = iwork + n, *nextcol = iwork + 2 * n;
    idx npos = 0;
    *reg_count = 0;

    Lp[0] = 0;
    for (idx i = 0; i < n; i++) Lp[i + 1] = Lp[i] + Lnz[i];
    for (idx i = 0; i < n; i++) {
        ymark[i] = 0; yval[i] = 0.0; D[i] = 0.0; nextcol[i] = Lp[i];
    }

    if (!logical && n > 0) {
        /* first pivot special case, qdldl.rs:516-534 */
        D[0] = Ax[0];
        if (reg_enable) {
            double s = (double)Dsigns[0];
            if (D[0] * s < reg_eps) { D[0] = reg_delta * s; (*reg_count)++; }
        }
        if (D[0] == 0.0) return OQ_ERR_ZERO_PIVOT;
        if (D[0] > 0.0) npos++;
        Dinv[0] = 1.0 / D[0];
    }

    for (idx k = 1; k < n; k++) {
        idx nnzy = 0;
        /* pattern of row k of L: walk the etree from each above-diagonal
           entry of column k, collecting unvisited path segments */
        for (idx p = Ap[k]; p < Ap[k + 1]; p++) {
            idx b = Ai[p];
            if (b == k) { D[k] = Ax[p]; continue; }
            yval[b] = Ax[p];
            if (!ymark[b]) {
                ymark[b] = 1;
                ebuf[0] = b;
                idx ne = 1;
                idx nx = etree[b];
                while (nx != OQ_UNKNOWN && nx < k) {
                    if (ymark[nx]) break;
                    ymark[nx] = 1;
                    ebuf[ne++] = nx;
                    nx = etree[nx];
                }
                while (ne) yidx[nnzy++] = ebuf[--ne];
            }
        }
        /* numeric sparse triangular solve, in reverse list order */
        for (idx t = nnzy - 1; t >= 0; t--) {
            idx c = yidx[t];
            idx dst = nextcol[c];
            if (!logical) {
                double yc = yval[c];
                for (idx q = Lp[c]; q < dst; q++) yval[Li[q]] -= Lx[q] * yc;
                double l = yc * Dinv[c];
                Lx[dst] = l;
                D[k] -= yc * l;
            }
            Li[dst] = k;
            nextcol[c] = dst + 1;
            yval[c] = 0.0;
            ymark[c] = 0;
        }
        if (!logical) {
            if (reg_enable) {
                double s = (double)Dsigns[k];
                if (D[k] * s < reg_eps) { D[k] = reg_delta * s; (*reg_count)++; }
            }
            if (D[k] == 0.0) return OQ_ERR_ZERO_PIVOT;
            if (D[k] > 0.0) npos++;
            Dinv[k] = 1.0 / D[k];
        }
    }
    *pos_inertia = npos;
    return OQ_OK;
}

/* ---- triangular solves --------------------------------------------------- */

void oq_lsolve(idx n, const idx *Lp, const idx *Li, const double *Lx, double *x)
{
    for (idx i = 0; i < n; i++) {
        double xi = x[i];
        for (idx q = Lp[i]; q < Lp[i + 1]; q++) x[Li[q]] -= Lx[q] * xi;
    }
}

void oq_ltsolve(idx n, const idx *Lp, const idx *Li, const double *Lx, double *x)
{
    for (idx i = n - 1; i >= 0; i--) {
        double s = 0.0;
        for (idx q = Lp[i]; q < Lp[i + 1]; q++) s += Lx[q] * x[Li[q]];
        x[i] -= s;
    }
}

void oq_dltsolve(idx n, const idx *Lp, const idx *Li, const double *Lx,
                 const double *Dinv, double *x)
{
    for (idx i = n - 1; i >= 0; i--) {
        double s = 0.0;
        for (idx q = Lp[i]; q < Lp[i + 1]; q++) s += Lx[q] * x[Li[q]];
        x[i] *= Dinv[i];
        x[i] -= s;
    }
}

void oq_solve_factors(idx n, const idx *Lp, const idx *Li, const double *Lx,
                      const double *Dinv, double *b)
{
    oq_lsolve(n, Lp, Li, Lx, b);
    oq_dltsolve(n, Lp, Li, Lx, Dinv, b);
}

/* ---- object API ---------------------------------------------------------- */

static int check_structure(idx nrows, idx ncols, const idx *Ap, const idx *Ai)
{
    if (nrows != ncols) return OQ_ERR_DIM;
    for (idx j = 0; j < ncols; j++)
        for (idx p = Ap[j]; p < Ap[j + 1]; p++)
            if (Ai[p] > j) return OQ_ERR_NOT_TRIU;
    for (idx j = 0; j < ncols; j++)
        if (!(Ap[j] < Ap[j + 1])) return OQ_ERR_EMPTYCOL;
    return OQ_OK;
}

void oq_free(oq_t *f)
{
    if (!f) return;
    free(f->perm); free(f->iperm); free(f->Ap); free(f->Ai); free(f->Ax);
    free(f->AtoPAPt); free(f->Lp); free(f->Li); free(f->Lx); free(f->D);
    free(f->Dinv); free(f->etree); free(f->Lnz); free(f->iwork);
    free(f->bwork); free(f->fwork); free(f->Dsigns); free(f);
}

/* perm must be supplied (no AMD in the oracle).  Dsigns may be NULL (=> +1).
   `logical` != 0 reproduces the adapter's allocate-only construction
   (ldlsolvers/qdldl.rs:33-46). */
int oq_new(oq_t **out, idx nrows, idx ncols, const idx *Ap, const idx *Ai,
           const double *Ax, const idx *perm, const int8_t *Dsigns, int logical,
           int reg_enable, double reg_eps, double reg_delta)
{
    *out = NULL;
    int rc = check_structure(nrows, ncols, Ap, Ai);
    if (rc) return rc;
    idx n = ncols, nnz = Ap[n];
    size_t sn = (size_t)(n > 0 ? n : 1);
    oq_t *f = (oq_t *)calloc(1, sizeof(oq_t));
    f->n = n; f->nnzA = nnz;
    f->perm = (idx *)malloc(sn * sizeof(idx));
    f->iperm = (idx *)malloc(sn * sizeof(idx));
    memcpy(f->perm, perm, (size_t)n * sizeof(idx));
    rc = oq_invperm(n, perm, f->iperm);
    if (rc) { oq_free(f); return rc; }

    f->Ap = (idx *)malloc((sn + 1) * sizeof(idx));
    f->Ai = (idx *)malloc((size_t)(nnz + 1) * sizeof(idx));
    f->Ax = (double *)malloc((size_t)(nnz + 1) * sizeof(double));
    f->AtoPAPt = (idx *)malloc((size_t)(nnz + 1) * sizeof(idx));
    oq_permute_symmetric(n, Ap, Ai, Ax, f->iperm, f->Ap, f->Ai, f->Ax, f->AtoPAPt);

    f->Dsigns = (int8_t *)malloc(sn);
    for (idx i = 0; i < n; i++) f->Dsigns[i] = Dsigns ? Dsigns[perm[i]] : 1;
    f->regularize_enable = reg_enable;
    f->regularize_eps = reg_eps; f->regularize_delta = reg_delta;

    f->etree = (idx *)malloc(sn * sizeof(idx));
    f->Lnz = (idx *)malloc(sn * sizeof(idx));
    f->iwork = (idx *)malloc(3 * sn * sizeof(idx));
    f->bwork = (unsigned char *)malloc(sn);
    f->fwork = (double *)malloc(sn * sizeof(double));
    oq_etree(n, f->Ap, f->Ai, f->iwork, f->Lnz, f->etree);

    idx sumLnz = 0;
    for (idx i = 0; i < n; i++) sumLnz += f->Lnz[i];
    f->nnzL = sumLnz;
    f->Lp = (idx *)malloc((sn + 1) * sizeof(idx));
    f->Li = (idx *)malloc((size_t)(sumLnz + 1) * sizeof(idx));
    f->Lx = (double *)malloc((size_t)(sumLnz + 1) * sizeof(double));
    f->D = (double *)calloc(sn, sizeof(double));
    f->Dinv = (double *)calloc(sn, sizeof(double));
    f->is_symbolic = logical;
    if (logical) {
        for (idx i = 0; i < sumLnz; i++) f->Lx[i] = 1.0;
        for (idx i = 0; i < n; i++) { f->D[i] = 1.0; f->Dinv[i] = 1.0; }
    }
    rc = oq_factor(n, f->Ap, f->Ai, f->Ax, f->Lp, f->Li, f->Lx, f->D, f->Dinv,
                   f->Lnz, f->etree, f->bwork, f->iwork, f->fwork, logical,
                   f->Dsigns, reg_enable, reg_eps, reg_delta,
                   &f->regularize_count, &f->positive_inertia);
    if (rc) { oq_free(f); return rc; }
    *out = f;
    return OQ_OK;
}

int oq_refactor(oq_t *f)
{
    f->is_symbolic = 0;
    return oq_factor(f->n, f->Ap, f->Ai, f->Ax, f->Lp, f->Li, f->Lx, f->D,
                     f->Dinv, f->Lnz, f->etree, f->bwork, f->iwork, f->fwork, 0,
                     f->Dsigns, f->regularize_enable, f->regularize_eps,
                     f->regularize_delta, &f->regularize_count,
                     &f->positive_inertia);
}

/* in-place solve; returns -1 if only a logical factorisation exists
   (the reference panics there, qdldl.rs:118) */
int oq_solve(oq_t *f, double *b)
{
    if (f->is_symbolic) return -1;
    double *tmp = f->fwork;
    oq_permute(f->n, tmp, b, f->perm);
    oq_solve_factors(f->n, f->Lp, f->Li, f->Lx, f->Dinv, tmp);
    oq_ipermute(f->n, b, tmp, f->perm);
    return 0;
}

void oq_update_values(oq_t *f, const idx *index, const double *values, idx len)
{ for (idx i = 0; i < len; i++) f->Ax[f->AtoPAPt[index[i]]] = values[i]; }

void oq_scale_values(oq_t *f, const idx *index, idx len, double scale)
{ for (idx i = 0; i < len; i++) f->Ax[f->AtoPAPt[index[i]]] *= scale; }

void oq_offset_values(oq_t *f, const idx *index, idx len, double offset,
                      const int8_t *signs)
{
    for (idx i = 0; i < len; i++) {
        if (signs[i] > 0) f->Ax[f->AtoPAPt[index[i]]] += offset;
        else if (signs[i] < 0) f->Ax[f->AtoPAPt[index[i]]] -= offset;
    }
}

/* adapter-level refactor result: all Dinv finite (ldlsolvers/qdldl.rs:99-106) */
int oq_dinv_is_finite(const oq_t *f)
{
    for (idx i = 0; i < f->n; i++) if (!isfinite(f->Dinv[i])) return 0;
    return 1;
}

/* accessors for the tests */
idx oq_n(const oq_t *f) { return f->n; }
idx oq_nnzA(const oq_t *f) { return f->nnzA; }
idx oq_nnzL(const oq_t *f) { return f->nnzL; }
idx oq_regularize_count(const oq_t *f) { return f->regularize_count; }
idx oq_positive_inertia(const oq_t *f) { return f->positive_inertia; }
const idx *oq_Lp(const oq_t *f) { return f->Lp; }
const idx *oq_Li(const oq_t *f) { return f->Li; }
const double *oq_Lx(const oq_t *f) { return f->Lx; }
const double *oq_D(const oq_t *f) { return f->D; }
const double *oq_Dinv(const oq_t *f) { return f->Dinv; }
const idx *oq_etree_ptr(const oq_t *f) { return f->etree; }
const idx *oq_Lnz(const oq_t *f) { return f->Lnz; }
const idx *oq_permA_colptr(const oq_t *f) { return f->Ap; }
const idx *oq_permA_rowval(const oq_t *f) { return f->Ai; }
const double *oq_permA_nzval(const oq_t *f) { return f->Ax; }
const idx *oq_AtoPAPt(const oq_t *f) { return f->AtoPAPt; }
