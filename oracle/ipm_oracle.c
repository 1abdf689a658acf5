/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY.  Not product code.
 *
 * CPU restatement (plain C, single thread, f64) of the callers of the LDL hot
 * path: KKT assembly + index maps, the DirectLDLKKTSolver (static
 * regularisation, refactor, iterative refinement), the Zero / Nonnegative /
 * second-order cones, Ruiz equilibration and the interior-point main loop.
 * It exists so that "same solver status / same iteration count / same
 * solution" can be checked between the reference algorithm on the CPU and the
 * CUDA backend, and as the timed CPU baseline.
 *
 * Pinned against the reference's end-to-end known answers
 * (tests/basic_qp.rs, basic_lp.rs, basic_socp.rs, basic_eq_constrained.rs,
 * basic_unconstrained.rs, presolve.rs data, examples/data/hs35.json) and the
 * exact KKT patterns of kkt_assembly.rs:185-355 in tests/test_oracle_ipm.py.
 * The reference holds NO unit-level known answers for cone numerics
 * (SURVEY.md section 4): those are pinned end-to-end only.
 *
 * Not restated (documented gaps): chordal decomposition.
 * The inf-bound presolve (presolver.rs: nonnegative rows with b beyond the
 * infinity bound are dropped, the solution is expanded again) is restated in
 * oipm_new_ex / oipm_get_solution and pinned on tests/presolve.rs.
 * The exponential and 3-D power cones (nonsymmetric path: unit initialisation,
 * dual / primal-dual scaling, third-order correction, backtracking step
 * length, barrier line search, strategy checkpoints) live in nonsym_oracle.h
 * and are pinned on tests/basic_expcone.rs, basic_powcone.rs, mixed_conic.rs;
 * the generalised power cone (rank-3 sparse expansion, dual scaling only)
 * likewise, pinned on tests/basic_genpowcone.rs.
 * The LDL ordering is passed in (see qdldl_oracle.c header).
 *
 * Function -> reference map (all under /root/reference/src)
 *   collapse_cones               solver/core/cones/supportedcone.rs:105-161
 *   kkt_assemble                 .../quasidef/kkt_assembly.rs:20-183, datamaps.rs:112-221,
 *                                algebra/csc/utils.rs:16-307
 *   fill_signs                   .../quasidef/directldlkktsolver.rs:392-405
 *   kkt_update / regularize      directldlkktsolver.rs:134-158, 217-264, 324-329
 *   kkt_solve / refine           directldlkktsolver.rs:168-189, 266-347
 *   symv_triu                    algebra/csc/matrix_math.rs:178-208
 *   quad_form_triu               algebra/csc/matrix_math.rs:212-257
 *   nn_* / zero_* / soc_*        solver/core/cones/{nonnegativecone,zerocone,socone}.rs
 *   combined_ds_shift_symmetric  solver/core/cones/symmetric_common.rs:53-84
 *   equilibrate                  solver/implementations/default/problemdata.rs:229-349
 *   residuals_update             .../default/residuals.rs:69-111
 *   kktsystem_*                  .../default/kktsystem.rs:108-278
 *   variables_*                  .../default/variables.rs:62-285
 *   info_update/check_*          .../default/info.rs:112-389
 *   oipm_solve                   solver/core/solver.rs:242-465, 525-665
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef int64_t idx;
static double vnorm(const double *a, idx n);
/* ORACLE_JITTER=<seed>: multiply the outputs of the nonsymmetric cones' arithmetic by (1 + k eps), |k| <= ORACLE_JITTER_ULP
   (default 4), to
   probe how sensitive the discrete decisions of the driver (backtracking counts, strategy switches, iteration
   count) are to last-bit differences such as a GPU's libm against glibc (tests/test_oracle_nonsym.py).  Off unless
   the variable is set. */
static int jitter_on = -1;
static uint64_t jitter_state = 0;
static double jit(double x)
{
    if (jitter_on < 0) { const char *e = getenv("ORACLE_JITTER"); jitter_on = e ? 1 : 0; jitter_state = e ? (uint64_t)atoll(e) * 2654435761u + 88172645463325252ull : 0; }
    if (!jitter_on) return x;
    jitter_state ^= jitter_state << 13; jitter_state ^= jitter_state >> 7; jitter_state ^= jitter_state << 17;
    const int64_t amp = getenv("ORACLE_JITTER_ULP") ? atoll(getenv("ORACLE_JITTER_ULP")) : 4;
    const double k = (double)((int64_t)(jitter_state % (uint64_t)(2 * amp + 1)) - amp);
    return x * (1.0 + k * 2.220446049250313e-16);
}
#include "nonsym_oracle.h"

/* from qdldl_oracle.c */
typedef struct oq_s oq_t;
int oq_new(oq_t **out, idx nrows, idx ncols, const idx *Ap, const idx *Ai, const double *Ax,
           const idx *perm, const int8_t *Dsigns, int logical, int reg_enable, double reg_eps,
           double reg_delta);
void oq_free(oq_t *f);
int oq_refactor(oq_t *f);
int oq_solve(oq_t *f, double *b);
void oq_update_values(oq_t *f, const idx *index, const double *values, idx len);
void oq_scale_values(oq_t *f, const idx *index, idx len, double scale);
int oq_dinv_is_finite(const oq_t *f);
idx oq_nnzL(const oq_t *f);
idx oq_regularize_count(const oq_t *f);

/* src/utils/infbounds.rs:10-40: the process-wide "infinite" bound, default 1e20 */
static double g_infinity = 1e20;
double oipm_get_infinity(void) { return g_infinity; }
void oipm_set_infinity(double v) { g_infinity = v; }
void oipm_default_infinity(void) { g_infinity = 1e20; }

enum { CONE_ZERO = 0, CONE_NONNEG = 1, CONE_SOC = 2, CONE_PSD = 3, CONE_EXP = 4, CONE_POW = 5, CONE_GENPOW = 6 };
enum { SCALING_PRIMAL_DUAL = 0, SCALING_DUAL = 1 };
enum { ST_UNSOLVED = 0, ST_SOLVED, ST_PRIMAL_INFEASIBLE, ST_DUAL_INFEASIBLE, ST_ALMOST_SOLVED,
       ST_ALMOST_PRIMAL_INFEASIBLE, ST_ALMOST_DUAL_INFEASIBLE, ST_MAX_ITERATIONS, ST_MAX_TIME,
       ST_NUMERICAL_ERROR, ST_INSUFFICIENT_PROGRESS };

/* mirrors the fields of DefaultSettings the path reads (default/settings.rs:30-193) */
typedef struct {
    int32_t max_iter;
    double time_limit;
    double max_step_fraction;
    double tol_gap_abs, tol_gap_rel, tol_feas, tol_infeas_abs, tol_infeas_rel, tol_ktratio;
    double reduced_tol_gap_abs, reduced_tol_gap_rel, reduced_tol_feas, reduced_tol_infeas_abs,
        reduced_tol_infeas_rel, reduced_tol_ktratio;
    int32_t equilibrate_enable, equilibrate_max_iter;
    double equilibrate_min_scaling, equilibrate_max_scaling;
    double min_terminate_step_length;
    int32_t static_regularization_enable;
    double static_regularization_constant, static_regularization_proportional;
    int32_t dynamic_regularization_enable;
    double dynamic_regularization_eps, dynamic_regularization_delta;
    int32_t iterative_refinement_enable;
    double iterative_refinement_reltol, iterative_refinement_abstol;
    int32_t iterative_refinement_max_iter;
    double iterative_refinement_stop_ratio;
    /* nonsymmetric cones only (settings.rs:114-124) */
    double linesearch_backtrack_step, min_switch_step_length;
    int32_t presolve_enable;
} oipm_settings;

void oipm_default_settings(oipm_settings *s)
{
    s->max_iter = 200; s->time_limit = INFINITY; s->max_step_fraction = 0.99;
    s->tol_gap_abs = 1e-8; s->tol_gap_rel = 1e-8; s->tol_feas = 1e-8;
    s->tol_infeas_abs = 1e-8; s->tol_infeas_rel = 1e-8; s->tol_ktratio = 1e-6;
    s->reduced_tol_gap_abs = 5e-5; s->reduced_tol_gap_rel = 5e-5; s->reduced_tol_feas = 1e-4;
    s->reduced_tol_infeas_abs = 5e-12; s->reduced_tol_infeas_rel = 5e-5; s->reduced_tol_ktratio = 1e-4;
    s->equilibrate_enable = 1; s->equilibrate_max_iter = 10;
    s->equilibrate_min_scaling = 1e-4; s->equilibrate_max_scaling = 1e4;
    s->min_terminate_step_length = 1e-4;
    s->static_regularization_enable = 1; s->static_regularization_constant = 1e-8;
    s->static_regularization_proportional = 2.220446049250313e-16 * 2.220446049250313e-16;
    s->dynamic_regularization_enable = 1; s->dynamic_regularization_eps = 1e-13;
    s->dynamic_regularization_delta = 2e-7;
    s->iterative_refinement_enable = 1; s->iterative_refinement_reltol = 1e-13;
    s->iterative_refinement_abstol = 1e-12; s->iterative_refinement_max_iter = 10;
    s->iterative_refinement_stop_ratio = 5.0;
    s->linesearch_backtrack_step = 0.8; s->min_switch_step_length = 0.1;
    s->presolve_enable = 1;
}

typedef struct {
    int32_t status, iterations;
    double cost_primal, cost_dual, res_primal, res_dual, res_primal_inf, res_dual_inf;
    double gap_abs, gap_rel, ktratio, mu, step_length, sigma;
    double solve_time, t_kkt_update, t_kkt_solve, t_scale_cones;
    int64_t n_refactor, n_ldl_solve, nnzL, nnzK;
} oipm_info;

typedef struct {
    int type; idx dim, off, boff, blen;
    double *w, *lam; double eta;
    int sparse; double *u, *v; double d;
    idx *map_u, *map_v; idx map_D[2];
    /* PSD triangle cone (psdtrianglecone.rs:12-60): matrix dimension and dense work data */
    idx psd_n; double *R, *Rinv, *lisqrt, *HsM, *W1, *W2, *W3, *wv;
    /* exponential / 3-D power cone state */
    ns3_t *ns;
    /* generalised power cone state */
    gp_t *gp;
} cone_t;

typedef struct { idx m, n; idx *colptr, *rowval; double *nzval; } csc;

typedef struct {
    idx n, m, p, N;
    idx mfull; char *keep;   /* presolve: rows of the user's problem kept in the reduced one (NULL = all) */
    csc P, A;              /* internal (scaled) copies; P is triu */
    double *q, *b;
    double normq, normb;
    idx ncones; cone_t *cones; idx degree; int all_symmetric, allows_primal_dual;
    /* equilibration */
    double *d, *dinv, *e, *einv, c;
    /* KKT */
    csc K; idx *map_P, *map_A, *map_Hs, *map_diagP, *map_diag_full; idx nHs;
    int8_t *dsigns; double *Hs;
    double *kx, *kb, *kw1, *kw2;  /* x, b, work1, work2 of DirectLDLKKTSolver */
    oq_t *ldl; idx *perm;
    double diagonal_regularizer;
    /* kktsystem */
    double *x1, *z1, *x2, *z2, *workx, *workz, *work_conic;
    /* variables: x s z tau kappa  (vars, step_lhs, step_rhs, prev) */
    double *vx, *vs, *vz, vtau, vkap;
    double *lx, *ls, *lz, ltau, lkap;
    double *rx_, *rs_, *rz_, rtau_, rkap_;
    double *px, *ps, *pz, ptau, pkap;
    /* residuals */
    double *rx, *rz, rtau, *rx_inf, *rz_inf, *Px; double dot_qx, dot_bz, dot_sz, dot_xPx;
    oipm_settings set; oipm_info info;
    double prev_cost_primal, prev_cost_dual, prev_res_primal, prev_res_dual, prev_gap_abs, prev_gap_rel;
} oipm_t;

/* ---------------------------------------------------------------- vectors */
static double now_s(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }
static double *dvec(idx n) { return (double *)calloc((size_t)(n > 0 ? n : 1), sizeof(double)); }
static idx *ivec(idx n) { return (idx *)calloc((size_t)(n > 0 ? n : 1), sizeof(idx)); }
static double vdot(const double *a, const double *b, idx n) { double s = 0; for (idx i = 0; i < n; i++) s += a[i] * b[i]; return s; }
static double vnorm_inf(const double *a, idx n)
{ double o = 0; for (idx i = 0; i < n; i++) { if (isnan(a[i])) return NAN; double v = fabs(a[i]); if (v > o) o = v; } return o; }
/* overflow-safe 2-norm, algebra/vecmath.rs:206-226 */
typedef struct { double scale, sumsq; } sn_t;
static inline void sn_add(sn_t *s, double xi)
{
    if (xi == 0.0) return;
    double a = fabs(xi);
    if (s->scale < a) { double r = s->scale / a; s->sumsq = 1.0 + s->sumsq * r * r; s->scale = a; }
    else { double r = a / s->scale; s->sumsq += r * r; }
}
static double vnorm(const double *a, idx n) { sn_t s = {0.0, 1.0}; for (idx i = 0; i < n; i++) sn_add(&s, a[i]); return s.scale * sqrt(s.sumsq); }
static double vnorm_scaled(const double *a, const double *v, idx n) { sn_t s = {0.0, 1.0}; for (idx i = 0; i < n; i++) sn_add(&s, a[i] * v[i]); return s.scale * sqrt(s.sumsq); }
static double vmean(const double *a, idx n) { if (n == 0) return 0; double s = 0; for (idx i = 0; i < n; i++) s += a[i]; return s / (double)n; }
static int vfinite(const double *a, idx n) { for (idx i = 0; i < n; i++) if (!isfinite(a[i])) return 0; return 1; }
static double clipd(double x, double lo, double hi) { return x < lo ? lo : (x > hi ? hi : x); }

/* ------------------------------------------------------------ sparse ops */
static void csc_copy(csc *dst, idx m, idx n, const idx *cp, const idx *rv, const double *nz)
{
    idx nnz = cp[n];
    dst->m = m; dst->n = n; dst->colptr = ivec(n + 1); dst->rowval = ivec(nnz); dst->nzval = dvec(nnz);
    memcpy(dst->colptr, cp, (size_t)(n + 1) * sizeof(idx));
    memcpy(dst->rowval, rv, (size_t)nnz * sizeof(idx));
    memcpy(dst->nzval, nz, (size_t)nnz * sizeof(double));
}
static void csc_free(csc *a) { free(a->colptr); free(a->rowval); free(a->nzval); }

/* y = a*K*x + b*y, K symmetric stored as one triangle */
static void symv_tri(const csc *A, double *y, const double *x, double a, double b)
{
    idx n = A->n;
    for (idx i = 0; i < n; i++) y[i] *= b;
    for (idx col = 0; col < n; col++) {
        double xc = x[col];
        for (idx p = A->colptr[col]; p < A->colptr[col + 1]; p++) {
            idx row = A->rowval[p]; double v = A->nzval[p];
            y[row] += a * v * xc;
            if (row != col) y[col] += a * v * x[row];
        }
    }
}
static double quad_form_triu(const csc *M, const double *y, const double *x)
{
    double out = 0;
    for (idx col = 0; col < M->n; col++) {
        double t1 = 0, t2 = 0;
        for (idx p = M->colptr[col]; p < M->colptr[col + 1]; p++) {
            idx row = M->rowval[p]; double v = M->nzval[p];
            if (row < col) { t1 += v * x[row]; t2 += v * y[row]; }
            else if (row == col) out += v * x[col] * y[col];
        }
        out += t1 * y[col] + t2 * x[col];
    }
    return out;
}
static void scale_by_b(double *y, idx n, double b)
{
    if (b == 0.0) { for (idx i = 0; i < n; i++) y[i] = 0.0; }
    else if (b == 1.0) { }
    else if (b == -1.0) { for (idx i = 0; i < n; i++) y[i] = -y[i]; }
    else for (idx i = 0; i < n; i++) y[i] *= b;
}
static void gemv_N(const csc *A, double *y, const double *x, double a, double b)
{
    scale_by_b(y, A->m, b);
    if (a == 0.0) return;
    for (idx j = 0; j < A->n; j++) {
        double xj = x[j];
        for (idx p = A->colptr[j]; p < A->colptr[j + 1]; p++) {
            if (a == 1.0) y[A->rowval[p]] += A->nzval[p] * xj;
            else if (a == -1.0) y[A->rowval[p]] -= A->nzval[p] * xj;
            else y[A->rowval[p]] += a * A->nzval[p] * xj;
        }
    }
}
static void gemv_T(const csc *A, double *y, const double *x, double a, double b)
{
    scale_by_b(y, A->n, b);
    if (a == 0.0) return;
    for (idx j = 0; j < A->n; j++) {
        double yj = y[j];
        for (idx p = A->colptr[j]; p < A->colptr[j + 1]; p++) {
            if (a == 1.0) yj += A->nzval[p] * x[A->rowval[p]];
            else if (a == -1.0) yj -= A->nzval[p] * x[A->rowval[p]];
            else yj += a * A->nzval[p] * x[A->rowval[p]];
        }
        y[j] = yj;
    }
}

/* ------------------------------------------------------------------ cones */
static double soc_residual(const double *z, idx n) { double t = vnorm(z + 1, n - 1); return (z[0] - t) * (z[0] + t); }
static double soc_sqrt_residual(const double *z, idx n) { double r = soc_residual(z, n); return r > 0 ? sqrt(r) : 0.0; }

static void soc_circ(double *x, const double *y, const double *z, idx n)
{
    double x0 = vdot(y, z, n); double y0 = y[0], z0 = z[0];
    for (idx i = 1; i < n; i++) x[i] = y0 * z[i] + z0 * y[i];
    x[0] = x0;
}
static void soc_mul_W(double *y, const double *x, double a, double b, const double *w, double eta, idx n)
{
    double zeta = vdot(w + 1, x + 1, n - 1);
    double c = x[0] + zeta / (1.0 + w[0]);
    y[0] = (a * eta) * (w[0] * x[0] + zeta) + b * y[0];
    for (idx i = 1; i < n; i++) y[i] = (a * eta * c) * w[i] + b * y[i];
    for (idx i = 1; i < n; i++) y[i] = (a * eta) * x[i] + y[i];
}
static void soc_mul_Winv(double *y, const double *x, double a, double b, const double *w, double eta, idx n)
{
    double zeta = vdot(w + 1, x + 1, n - 1);
    double c = -x[0] + zeta / (1.0 + w[0]);
    y[0] = (a / eta) * (w[0] * x[0] - zeta) + b * y[0];
    for (idx i = 1; i < n; i++) y[i] = (a / eta * c) * w[i] + b * y[i];
    for (idx i = 1; i < n; i++) y[i] = (a / eta) * x[i] + y[i];
}
static double soc_step_component(const double *x, const double *y, idx n, double amax)
{
    if (x[0] >= 0.0 && y[0] < 0.0) { double t = -x[0] / y[0]; if (t < amax) amax = t; }
    double a = soc_residual(y, n);
    double b = 2.0 * (x[0] * y[0] - vdot(x + 1, y + 1, n - 1));
    double c = soc_residual(x, n); if (c < 0.0) c = 0.0;
    double d = b * b - 4.0 * a * c;
    if ((a > 0.0 && b > 0.0) || d < 0.0) return amax;
    else if (a == 0.0) return amax;
    else if (c == 0.0) return a >= 0.0 ? amax : 0.0;
    double t = b >= 0.0 ? (-b - sqrt(d)) : (-b + sqrt(d));
    double r1 = (2.0 * c) / t, r2 = t / (2.0 * a);
    if (r1 < 0.0) r1 = INFINITY;
    if (r2 < 0.0) r2 = INFINITY;
    double r = r1 < r2 ? r1 : r2;
    return amax < r ? amax : r;
}


/* ------------------------------------------------ PSD triangle cone helpers
 * Dense column-major n x n work; restates psdtrianglecone.rs with the LAPACK
 * calls (dpotrf, dgesdd, dsyevr -- third-party, not under /root/reference)
 * replaced by textbook algorithms: Cholesky, one-sided Jacobi SVD (Hestenes),
 * cyclic Jacobi eigenvalues.  Results are pinned end-to-end on tests/basic_sdp.rs
 * and against numpy's LAPACK in tests/test_oracle_psd.py. */
#define MAT(A, n, i, j) (A)[(size_t)(j) * (size_t)(n) + (size_t)(i)]
static idx tri_index(idx k) { return (k * (k + 3)) >> 1; }

static void svec_to_mat(double *M, idx n, const double *x)
{   /* dense/matrix_math.rs:165-183 */
    idx t = 0;
    for (idx col = 0; col < n; col++) for (idx row = 0; row <= col; row++) {
        if (row == col) MAT(M, n, row, col) = x[t];
        else { MAT(M, n, row, col) = x[t] * 0.70710678118654752440; MAT(M, n, col, row) = x[t] * 0.70710678118654752440; }
        t++;
    }
}
static void mat_to_svec(double *x, const double *M, idx n)
{   /* dense/matrix_math.rs:186-203 */
    idx t = 0;
    for (idx col = 0; col < n; col++) for (idx row = 0; row <= col; row++) {
        x[t] = (row == col) ? MAT(M, n, row, col) : (MAT(M, n, row, col) + MAT(M, n, col, row)) * 0.70710678118654752440;
        t++;
    }
}
/* lower Cholesky factor of the symmetric A into L (upper part zero); -1 if not positive definite */
static int chol_lower(double *L, const double *A, idx n)
{
    for (idx j = 0; j < n; j++) for (idx i = 0; i < n; i++) MAT(L, n, i, j) = 0.0;
    for (idx j = 0; j < n; j++) {
        double d = MAT(A, n, j, j);
        for (idx k = 0; k < j; k++) d -= MAT(L, n, j, k) * MAT(L, n, j, k);
        if (!(d > 0.0)) return -1;
        d = sqrt(d);
        MAT(L, n, j, j) = d;
        for (idx i = j + 1; i < n; i++) {
            double v = MAT(A, n, i, j);
            for (idx k = 0; k < j; k++) v -= MAT(L, n, i, k) * MAT(L, n, j, k);
            MAT(L, n, i, j) = v / d;
        }
    }
    return 0;
}
/* eigenvalues of a symmetric matrix (destroyed) by cyclic Jacobi */
static void sym_eigvals(double *A, idx n, double *lam)
{
    for (int sweep = 0; sweep < 60; sweep++) {
        double off = 0.0, dg = 0.0;
        for (idx j = 0; j < n; j++) for (idx i = 0; i < n; i++) { double v = MAT(A, n, i, j); if (i == j) dg += v * v; else off += v * v; }
        if (off <= 1e-32 * (dg + off) || off == 0.0) break;
        for (idx p = 0; p < n - 1; p++) for (idx q = p + 1; q < n; q++) {
            double apq = MAT(A, n, p, q);
            if (apq == 0.0) continue;
            double theta = (MAT(A, n, q, q) - MAT(A, n, p, p)) / (2.0 * apq);
            double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
            double c = 1.0 / sqrt(t * t + 1.0), sn = t * c;
            for (idx k = 0; k < n; k++) { double akp = MAT(A, n, k, p), akq = MAT(A, n, k, q); MAT(A, n, k, p) = c * akp - sn * akq; MAT(A, n, k, q) = sn * akp + c * akq; }
            for (idx k = 0; k < n; k++) { double apk = MAT(A, n, p, k), aqk = MAT(A, n, q, k); MAT(A, n, p, k) = c * apk - sn * aqk; MAT(A, n, q, k) = sn * apk + c * aqk; }
        }
    }
    for (idx i = 0; i < n; i++) lam[i] = MAT(A, n, i, i);
}
/* M = U diag(s) V^T by one-sided Jacobi; singular values sorted descending like LAPACK */
static void jacobi_svd(const double *M, idx n, double *U, double *sv, double *V)
{
    memcpy(U, M, (size_t)(n * n) * sizeof(double));
    for (idx j = 0; j < n; j++) for (idx i = 0; i < n; i++) MAT(V, n, i, j) = (i == j) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 60; sweep++) {
        int rotated = 0;
        for (idx p = 0; p < n - 1; p++) for (idx q = p + 1; q < n; q++) {
            double a = 0, b = 0, g = 0;
            for (idx k = 0; k < n; k++) { double up = MAT(U, n, k, p), uq = MAT(U, n, k, q); a += up * up; b += uq * uq; g += up * uq; }
            if (fabs(g) <= 1e-15 * sqrt(a * b) || g == 0.0) continue;
            rotated = 1;
            double zeta = (b - a) / (2.0 * g);
            double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
            double c = 1.0 / sqrt(1.0 + t * t), sn = c * t;
            for (idx k = 0; k < n; k++) { double up = MAT(U, n, k, p), uq = MAT(U, n, k, q); MAT(U, n, k, p) = c * up - sn * uq; MAT(U, n, k, q) = sn * up + c * uq; }
            for (idx k = 0; k < n; k++) { double vp = MAT(V, n, k, p), vq = MAT(V, n, k, q); MAT(V, n, k, p) = c * vp - sn * vq; MAT(V, n, k, q) = sn * vp + c * vq; }
        }
        if (!rotated) break;
    }
    for (idx j = 0; j < n; j++) {
        double nn = 0; for (idx k = 0; k < n; k++) nn += MAT(U, n, k, j) * MAT(U, n, k, j);
        sv[j] = sqrt(nn);
        if (sv[j] > 0) for (idx k = 0; k < n; k++) MAT(U, n, k, j) /= sv[j];
    }
    for (idx a = 0; a < n - 1; a++) {       /* selection sort, descending */
        idx m = a; for (idx b = a + 1; b < n; b++) if (sv[b] > sv[m]) m = b;
        if (m != a) {
            double t = sv[a]; sv[a] = sv[m]; sv[m] = t;
            for (idx k = 0; k < n; k++) { t = MAT(U, n, k, a); MAT(U, n, k, a) = MAT(U, n, k, m); MAT(U, n, k, m) = t; t = MAT(V, n, k, a); MAT(V, n, k, a) = MAT(V, n, k, m); MAT(V, n, k, m) = t; }
        }
    }
}
/* C = alpha * op(A) op(B) + beta * C, all n x n */
static void gemm_nn(double *C, const double *A, int ta, const double *B, int tb, idx n, double alpha, double beta)
{
    for (idx j = 0; j < n; j++) for (idx i = 0; i < n; i++) {
        double acc = 0.0;
        for (idx k = 0; k < n; k++) acc += (ta ? MAT(A, n, k, i) : MAT(A, n, i, k)) * (tb ? MAT(B, n, j, k) : MAT(B, n, k, j));
        MAT(C, n, i, j) = alpha * acc + (beta == 0.0 ? 0.0 : beta * MAT(C, n, i, j));
    }
}
/* psdtrianglecone.rs:364-396 */
static void psd_mul_Wx(const cone_t *c, int transpose, double *y, const double *x, double alpha, double beta, const double *Rx)
{
    idx n = c->psd_n; double *X = c->W1, *Y = c->W2, *tmp = c->W3;
    svec_to_mat(X, n, x); svec_to_mat(Y, n, y);
    if (transpose) { gemm_nn(tmp, X, 0, Rx, 1, n, 1.0, 0.0); gemm_nn(Y, Rx, 0, tmp, 0, n, alpha, beta); }
    else { gemm_nn(tmp, Rx, 1, X, 0, n, 1.0, 0.0); gemm_nn(Y, tmp, 0, Rx, 0, n, alpha, beta); }
    mat_to_svec(y, Y, n);
}
/* psdtrianglecone.rs:467-509 : upper triangle of the symmetric Kronecker product A (x)_s A */
static void psd_skron(double *out, idx N, const double *A, idx n)
{
    const double sqrt2 = 1.4142135623730951;
    idx col = 0;
    for (idx l = 0; l < n; l++) for (idx k = 0; k <= l; k++) {
        idx row = 0; int kl_eq = (k == l);
        for (idx j = 0; j < n && row <= col; j++) {
            double Ajl = MAT(A, n, j, l), Ajk = MAT(A, n, j, k);
            for (idx i = 0; i <= j; i++) {
                if (row > col) break;
                int ij_eq = (i == j);
                double v;
                if (!ij_eq && !kl_eq) v = MAT(A, n, i, k) * Ajl + MAT(A, n, i, l) * Ajk;
                else if (ij_eq && !kl_eq) v = sqrt2 * Ajl * Ajk;
                else if (!ij_eq && kl_eq) v = sqrt2 * MAT(A, n, i, l) * Ajk;
                else v = Ajl * Ajl;
                MAT(out, N, row, col) = v;
                row++;
            }
        }
        col++;
    }
}
static int psd_update_scaling(cone_t *c, const double *s, const double *z)
{   /* psdtrianglecone.rs:144-204 */
    idx n = c->psd_n; if (n == 0) return 1;
    double *S = c->W1, *Z = c->W2;
    double *L1 = dvec(n * n), *L2 = dvec(n * n), *M = dvec(n * n), *U = dvec(n * n), *V = dvec(n * n), *sv = dvec(n);
    int ok = 1;
    svec_to_mat(S, n, s); svec_to_mat(Z, n, z);
    if (chol_lower(L1, S, n) || chol_lower(L2, Z, n)) ok = 0;
    if (ok) {
        gemm_nn(M, L2, 1, L1, 0, n, 1.0, 0.0);           /* L2' L1 */
        jacobi_svd(M, n, U, sv, V);
        for (idx i = 0; i < n; i++) { c->lam[i] = sv[i]; c->lisqrt[i] = 1.0 / sqrt(sv[i]); }
        gemm_nn(c->R, L1, 0, V, 0, n, 1.0, 0.0);          /* R = L1 V Lambda^-1/2 */
        for (idx j = 0; j < n; j++) for (idx i = 0; i < n; i++) MAT(c->R, n, i, j) *= c->lisqrt[j];
        gemm_nn(c->Rinv, U, 1, L2, 1, n, 1.0, 0.0);       /* Rinv = Lambda^-1/2 U' L2' */
        for (idx j = 0; j < n; j++) for (idx i = 0; i < n; i++) MAT(c->Rinv, n, i, j) *= c->lisqrt[i];
        gemm_nn(c->W1, c->R, 0, c->R, 1, n, 1.0, 0.0);    /* R R' */
        psd_skron(c->HsM, c->dim, c->W1, n);
    }
    free(L1); free(L2); free(M); free(U); free(V); free(sv);
    return ok;
}
static double psd_step_component(cone_t *c, const double *d, double amax)
{   /* psdtrianglecone.rs:437-463 */
    idx n = c->psd_n; if (n == 0) return amax;
    double *Wk = c->W1;
    svec_to_mat(Wk, n, d);
    for (idx j = 0; j < n; j++) for (idx i = 0; i < n; i++) MAT(Wk, n, i, j) *= c->lisqrt[i] * c->lisqrt[j];
    double *ev = dvec(n);
    sym_eigvals(Wk, n, ev);
    double g = INFINITY; for (idx i = 0; i < n; i++) if (ev[i] < g) g = ev[i];
    free(ev);
    if (g < 0.0) { double t = -(1.0 / g); return t < amax ? t : amax; }
    return amax;
}
static void psd_circ(cone_t *c, double *x, const double *y, const double *z)
{   /* psdtrianglecone.rs:406-420: X = (YZ + ZY)/2 */
    idx n = c->psd_n; double *Y = c->W1, *Z = c->W2, *X = c->W3;
    svec_to_mat(Y, n, y); svec_to_mat(Z, n, z);
    for (idx j = 0; j < n; j++) for (idx i = 0; i < n; i++) {
        double acc = 0.0;
        for (idx k = 0; k < n; k++) acc += MAT(Y, n, i, k) * MAT(Z, n, k, j) + MAT(Z, n, i, k) * MAT(Y, n, k, j);
        MAT(X, n, i, j) = 0.5 * acc;
    }
    mat_to_svec(x, X, n);
}
static void psd_lambda_inv_circ(cone_t *c, double *x, const double *z)
{   /* psdtrianglecone.rs:317-332 */
    idx n = c->psd_n; double *X = c->W1, *Z = c->W2;
    svec_to_mat(Z, n, z);
    for (idx i = 0; i < n; i++) for (idx j = 0; j < n; j++) MAT(X, n, i, j) = (2.0 * MAT(Z, n, i, j)) / (c->lam[i] + c->lam[j]);
    mat_to_svec(x, X, n);
}

static int cone_is_sparse(const cone_t *c) { return c->type == CONE_SOC && c->sparse; }
/* number of extra KKT columns of a sparse-expandable cone (datamaps.rs:139-141, 245-247) */
static idx cone_pdim(const cone_t *c) { return cone_is_sparse(c) ? 2 : (c->type == CONE_GENPOW ? 3 : 0); }
static int cone_Hs_diag(const cone_t *c) { return c->type == CONE_ZERO || c->type == CONE_NONNEG || (c->type == CONE_SOC && c->sparse) || c->type == CONE_GENPOW; }
static int cone_is_ns3(const cone_t *c) { return c->type == CONE_EXP || c->type == CONE_POW; }
static int cone_is_nonsym(const cone_t *c) { return cone_is_ns3(c) || c->type == CONE_GENPOW; }
static idx cone_degree(const cone_t *c) { return c->type == CONE_ZERO ? 0 : (c->type == CONE_NONNEG ? c->dim : (c->type == CONE_PSD ? c->psd_n : (cone_is_ns3(c) ? 3 : (c->type == CONE_GENPOW ? c->gp->dim1 + 1 : 1)))); }

/* Cone::unit_initialization of every cone type (zerocone.rs:71-74, nonnegativecone.rs:68-71, socone.rs:114-119,
   psdtrianglecone.rs:131-136, expcone.rs:88-94, powcone.rs:79-87) */
static void cones_unit_initialization(oipm_t *S, double *z_, double *s_)
{
    for (idx k = 0; k < S->ncones; k++) {
        cone_t *c = &S->cones[k]; double *z = z_ + c->off, *s = s_ + c->off; idx n = c->dim;
        for (idx i = 0; i < n; i++) { s[i] = 0.0; z[i] = 0.0; }
        if (c->type == CONE_NONNEG) for (idx i = 0; i < n; i++) { z[i] = 1.0; s[i] = 1.0; }
        else if (c->type == CONE_SOC) { s[0] += 1.0; z[0] += 1.0; }
        else if (c->type == CONE_PSD) for (idx j = 0; j < c->psd_n; j++) { s[(j * (j + 3)) >> 1] += 1.0; z[(j * (j + 3)) >> 1] += 1.0; }
        else if (c->type == CONE_EXP) {
            s[0] = -1.051383945322714; s[1] = 0.556409619469370; s[2] = 1.258967884768947;
            z[0] = s[0]; z[1] = s[1]; z[2] = s[2];
        } else if (c->type == CONE_POW) {
            double a = c->ns->alpha;
            s[0] = sqrt(1.0 + a); s[1] = sqrt(1.0 + (1.0 - a)); s[2] = 0.0;
            z[0] = s[0]; z[1] = s[1]; z[2] = s[2];
        } else if (c->type == CONE_GENPOW) {      /* genpowcone.rs:132-141 */
            for (idx i = 0; i < c->gp->dim1; i++) s[i] = sqrt(1.0 + c->gp->alpha[i]);
            for (idx i = 0; i < n; i++) z[i] = s[i];
        }
    }
}

static void cones_set_identity(oipm_t *S)
{
    for (idx k = 0; k < S->ncones; k++) {
        cone_t *c = &S->cones[k];
        if (c->type == CONE_NONNEG) for (idx i = 0; i < c->dim; i++) c->w[i] = 1.0;
        else if (c->type == CONE_SOC) {
            for (idx i = 0; i < c->dim; i++) c->w[i] = 0.0;
            c->w[0] = 1.0; c->eta = 1.0;
            if (c->sparse) {
                c->d = 0.5;
                for (idx i = 0; i < c->dim; i++) { c->u[i] = 0.0; c->v[i] = 0.0; }
                c->u[0] = 0.70710678118654752440;
            }
        } else if (c->type == CONE_PSD) {
            idx n = c->psd_n, N = c->dim;
            for (idx j = 0; j < n; j++) for (idx i = 0; i < n; i++) { MAT(c->R, n, i, j) = (i == j); MAT(c->Rinv, n, i, j) = (i == j); }
            for (idx j = 0; j < N; j++) for (idx i = 0; i < N; i++) MAT(c->HsM, N, i, j) = (i == j);
        }
    }
}

static int cones_update_scaling(oipm_t *S, const double *s_, const double *z_, double mu, int strategy)
{
    for (idx k = 0; k < S->ncones; k++) {
        cone_t *c = &S->cones[k];
        const double *s = s_ + c->off, *z = z_ + c->off; idx n = c->dim;
        if (c->type == CONE_NONNEG) {
            for (idx i = 0; i < n; i++) { c->lam[i] = sqrt(s[i] * z[i]); c->w[i] = sqrt(s[i] / z[i]); }
        } else if (c->type == CONE_SOC) {
            double zscale = soc_sqrt_residual(z, n), sscale = soc_sqrt_residual(s, n);
            if (zscale == 0.0 || sscale == 0.0) return 0;
            c->eta = sqrt(sscale / zscale);
            double *w = c->w;
            double sinv = 1.0 / sscale;
            for (idx i = 0; i < n; i++) w[i] = s[i] * sinv;
            w[0] += z[0] / zscale;
            double mz = -(1.0 / zscale);
            for (idx i = 1; i < n; i++) w[i] = mz * z[i] + w[i];
            double wscale = soc_sqrt_residual(w, n);
            if (wscale == 0.0) return 0;
            double winv = 1.0 / wscale;
            for (idx i = 0; i < n; i++) w[i] *= winv;
            double w1sq = vdot(w + 1, w + 1, n - 1);
            w[0] = sqrt(1.0 + w1sq);
            double g = 0.5 * wscale;
            double *lam = c->lam;
            lam[0] = g;
            double ca = (g + z[0] / zscale) / sscale, cb = (g + s[0] / sscale) / zscale;
            for (idx i = 1; i < n; i++) lam[i] = ca * s[i] + cb * z[i];
            double den = 1.0 / (s[0] / sscale + z[0] / zscale + 2.0 * g);
            for (idx i = 1; i < n; i++) lam[i] *= den;
            double sq = sqrt(sscale * zscale);
            for (idx i = 0; i < n; i++) lam[i] *= sq;
            if (c->sparse) {
                double alpha = 2.0 * w[0];
                double wsq = w[0] * w[0] + w1sq;
                double wsqinv = 1.0 / wsq;
                c->d = 0.5 * wsqinv;
                double u0 = sqrt(wsq - c->d);
                double u1 = alpha / u0;
                double v1 = sqrt(2.0 * (2.0 + wsqinv) / (2.0 * wsq - wsqinv));
                c->u[0] = u0;
                for (idx i = 1; i < n; i++) c->u[i] = u1 * w[i];
                c->v[0] = 0.0;
                for (idx i = 1; i < n; i++) c->v[i] = v1 * w[i];
            }
        } else if (c->type == CONE_PSD) {
            if (!psd_update_scaling(c, s, z)) return 0;
        } else if (cone_is_ns3(c)) {
            /* expcone.rs:103-120, powcone.rs:96-113, nonsymmetric_common.rs:53-64 */
            ns3_t *K = c->ns;
            if (c->type == CONE_EXP) exp_update_dual_grad_H(K, z); else pow_update_dual_grad_H(K, z);
            if (strategy == SCALING_DUAL) ns3_use_dual_scaling(K, mu);
            else {
                double zt[3];
                if (c->type == CONE_EXP) exp_gradient_primal(s, zt); else pow_gradient_primal(s, K->alpha, zt);
                ns3_use_primal_dual_scaling(K, s, z, zt);
            }
            K->z[0] = z[0]; K->z[1] = z[1]; K->z[2] = z[2];
            for (int i = 0; i < 6; i++) { K->Hs[i] = jit(K->Hs[i]); K->H_dual[i] = jit(K->H_dual[i]); }
            for (int i = 0; i < 3; i++) K->grad[i] = jit(K->grad[i]);
        } else if (c->type == CONE_GENPOW) {      /* genpowcone.rs:149-163 */
            if (!gp_update_dual_grad_H(c->gp, z)) return 0;
            c->gp->mu = mu;
            for (idx i = 0; i < n; i++) { c->gp->z[i] = z[i]; c->gp->grad[i] = jit(c->gp->grad[i]); c->gp->p[i] = jit(c->gp->p[i]); }
            for (idx i = 0; i < c->gp->dim1; i++) { c->gp->q[i] = jit(c->gp->q[i]); c->gp->d1[i] = jit(c->gp->d1[i]); }
            for (idx i = 0; i < c->gp->dim2; i++) c->gp->r[i] = jit(c->gp->r[i]);
        }
    }
    return 1;
}

static void cones_get_Hs(const oipm_t *S, double *Hs)
{
    for (idx k = 0; k < S->ncones; k++) {
        const cone_t *c = &S->cones[k]; double *H = Hs + c->boff; idx n = c->dim;
        if (c->type == CONE_ZERO) for (idx i = 0; i < n; i++) H[i] = 0.0;
        else if (c->type == CONE_NONNEG) for (idx i = 0; i < n; i++) H[i] = c->w[i] * c->w[i];
        else if (c->type == CONE_SOC) {
            double e2 = c->eta * c->eta;
            if (c->sparse) { for (idx i = 0; i < n; i++) H[i] = e2; H[0] *= c->d; }
            else {
                const double *w = c->w;
                H[0] = (1.4142135623730951 * w[0] - 1.0) * (1.4142135623730951 * w[0] + 1.0);
                idx h = 1;
                for (idx col = 1; col < n; col++) {
                    for (idx row = 0; row <= col; row++) H[h++] = 2.0 * w[row] * w[col];
                    H[h - 1] += 1.0;
                }
                for (idx i = 0; i < c->blen; i++) H[i] *= e2;
            }
        } else if (c->type == CONE_PSD) {
            idx N = c->dim, t = 0;     /* pack_triu, dense/types.rs:187-201 */
            for (idx col = 0; col < N; col++) for (idx row = 0; row <= col; row++) H[t++] = MAT(c->HsM, N, row, col);
        } else if (cone_is_ns3(c)) for (int i = 0; i < 6; i++) H[i] = c->ns->Hs[i];   /* expcone.rs:126-129 */
        else if (c->type == CONE_GENPOW) {        /* genpowcone.rs:169-175: the diagonal D = [d1; d2] only */
            for (idx i = 0; i < c->gp->dim1; i++) H[i] = c->gp->mu * c->gp->d1[i];
            for (idx i = c->gp->dim1; i < n; i++) H[i] = c->gp->mu * c->gp->d2;
        }
    }
}

static void cones_mul_Hs(oipm_t *S, double *y_, const double *x_)
{
    for (idx k = 0; k < S->ncones; k++) {
        cone_t *c = &S->cones[k]; double *y = y_ + c->off; const double *x = x_ + c->off; idx n = c->dim;
        if (c->type == CONE_ZERO) for (idx i = 0; i < n; i++) y[i] = 0.0;
        else if (c->type == CONE_NONNEG) for (idx i = 0; i < n; i++) y[i] = c->w[i] * (c->w[i] * x[i]);
        else if (c->type == CONE_SOC) {
            double cc = vdot(c->w, x, n) * 2.0;
            for (idx i = 0; i < n; i++) y[i] = x[i];
            y[0] = -x[0];
            for (idx i = 0; i < n; i++) y[i] = cc * c->w[i] + y[i];
            double e2 = c->eta * c->eta;
            for (idx i = 0; i < n; i++) y[i] *= e2;
        } else if (c->type == CONE_PSD) {
            psd_mul_Wx(c, 0, c->wv, x, 1.0, 0.0, c->R);     /* work = W x  */
            psd_mul_Wx(c, 1, y, c->wv, 1.0, 0.0, c->R);     /* y = W' work */
        } else if (cone_is_ns3(c)) sym3_mul(c->ns->Hs, y, x);
        else if (c->type == CONE_GENPOW) gp_mul_Hs(c->gp, y, x);
    }
}

static void cones_affine_ds(const oipm_t *S, double *ds_, const double *s_)
{
    for (idx k = 0; k < S->ncones; k++) {
        const cone_t *c = &S->cones[k]; double *ds = ds_ + c->off; idx n = c->dim;
        if (c->type == CONE_ZERO) for (idx i = 0; i < n; i++) ds[i] = 0.0;
        else if (c->type == CONE_NONNEG) for (idx i = 0; i < n; i++) ds[i] = c->lam[i] * c->lam[i];
        else if (c->type == CONE_SOC) soc_circ(ds, c->lam, c->lam, n);
        else if (c->type == CONE_PSD) { for (idx i = 0; i < n; i++) ds[i] = 0.0; for (idx k = 0; k < c->psd_n; k++) ds[tri_index(k)] = c->lam[k] * c->lam[k]; }
        else if (cone_is_nonsym(c)) for (idx i = 0; i < n; i++) ds[i] = s_[c->off + i];   /* expcone.rs:135-137 */
    }
}

/* shift = (W^-T ds) o (W dz) - sigma*mu*e, steps modified in place */
static void cones_combined_ds_shift(oipm_t *S, double *shift_, double *sz_, double *ss_, double sigmamu)
{
    for (idx k = 0; k < S->ncones; k++) {
        cone_t *c = &S->cones[k]; idx n = c->dim;
        double *shift = shift_ + c->off, *sz = sz_ + c->off, *ss = ss_ + c->off;
        if (c->type == CONE_ZERO) { for (idx i = 0; i < n; i++) shift[i] = 0.0; continue; }
        if (cone_is_ns3(c)) {
            /* expcone.rs:139-148: third-order correction with the scaling-point z, inputs step_s, step_z */
            double eta[3] = {0, 0, 0};
            if (c->type == CONE_EXP) exp_higher_correction(c->ns, eta, ss, sz); else pow_higher_correction(c->ns, eta, ss, sz);
            for (int i = 0; i < 3; i++) shift[i] = jit(c->ns->grad[i] * sigmamu - eta[i]);
            continue;
        }
        if (c->type == CONE_GENPOW) { for (idx i = 0; i < n; i++) shift[i] = c->gp->grad[i] * sigmamu; continue; }   /* genpowcone.rs:208-213: no third-order term */
        double *tmp = shift;
        if (c->type == CONE_PSD) {
            memcpy(tmp, sz, (size_t)n * sizeof(double)); psd_mul_Wx(c, 0, sz, tmp, 1.0, 0.0, c->R);
            memcpy(tmp, ss, (size_t)n * sizeof(double)); psd_mul_Wx(c, 1, ss, tmp, 1.0, 0.0, c->Rinv);
            psd_circ(c, shift, ss, sz);
            for (idx k = 0; k < c->psd_n; k++) shift[tri_index(k)] += -sigmamu;
            continue;
        }
        memcpy(tmp, sz, (size_t)n * sizeof(double));
        if (c->type == CONE_NONNEG) for (idx i = 0; i < n; i++) sz[i] = 1.0 * (tmp[i] * c->w[i]) + 0.0 * sz[i];
        else soc_mul_W(sz, tmp, 1.0, 0.0, c->w, c->eta, n);
        memcpy(tmp, ss, (size_t)n * sizeof(double));
        if (c->type == CONE_NONNEG) for (idx i = 0; i < n; i++) ss[i] = 1.0 * (tmp[i] / c->w[i]) + 0.0 * ss[i];
        else soc_mul_Winv(ss, tmp, 1.0, 0.0, c->w, c->eta, n);
        if (c->type == CONE_NONNEG) { for (idx i = 0; i < n; i++) shift[i] = ss[i] * sz[i]; for (idx i = 0; i < n; i++) shift[i] += -sigmamu; }
        else { soc_circ(shift, ss, sz, n); shift[0] += -sigmamu; }
    }
}

static void cones_ds_from_dz_offset(oipm_t *S, double *out_, const double *ds_, const double *z_)
{
    for (idx k = 0; k < S->ncones; k++) {
        cone_t *c = &S->cones[k]; idx n = c->dim;
        double *out = out_ + c->off; const double *ds = ds_ + c->off, *z = z_ + c->off;
        if (c->type == CONE_ZERO) for (idx i = 0; i < n; i++) out[i] = 0.0;
        else if (c->type == CONE_NONNEG) for (idx i = 0; i < n; i++) out[i] = ds[i] / z[i];
        else if (c->type == CONE_PSD) { psd_lambda_inv_circ(c, c->wv, ds); psd_mul_Wx(c, 1, out, c->wv, 1.0, 0.0, c->R); }
        else if (cone_is_nonsym(c)) for (idx i = 0; i < n; i++) out[i] = ds[i];           /* expcone.rs:150-152 */
        else {
            double resz = soc_residual(z, n);
            double l1ds1 = vdot(c->lam + 1, ds + 1, n - 1), w1ds1 = vdot(c->w + 1, ds + 1, n - 1);
            for (idx i = 0; i < n; i++) out[i] = -z[i];
            out[0] = z[0];
            double cc = c->lam[0] * ds[0] - l1ds1;
            double sc = cc / resz;
            for (idx i = 0; i < n; i++) out[i] *= sc;
            out[0] += c->eta * w1ds1;
            for (idx i = 1; i < n; i++) out[i] += c->eta * (ds[i] + w1ds1 / (1.0 + c->w[0]) * c->w[i]);
            double li = 1.0 / c->lam[0];
            for (idx i = 0; i < n; i++) out[i] *= li;
        }
    }
}

/* nonsymmetric_common.rs:160-189 */
static double ns3_backtrack_search(const cone_t *c, const double *dq, const double *q, double a_init, double a_min,
                                   double step, int dual)
{
    double a = a_init, work3[3];
    double *work = c->type == CONE_GENPOW ? c->gp->work : work3;
    for (;;) {
        for (idx i = 0; i < c->dim; i++) work[i] = jit(1.0 * q[i] + a * dq[i]);
        int ok = c->type == CONE_GENPOW ? (dual ? gp_is_dual_feasible(c->gp, work) : gp_is_primal_feasible(c->gp, work))
               : c->type == CONE_EXP ? (dual ? exp_is_dual_feasible(work) : exp_is_primal_feasible(work))
                                     : (dual ? pow_is_dual_feasible(work, c->ns->alpha) : pow_is_primal_feasible(work, c->ns->alpha));
        if (ok) break;
        a *= step;
        if (a < a_min) { a = 0.0; break; }
    }
    return a;
}

/* compositecone.rs:289-332: symmetric cones first, then back off from a full step, then the nonsymmetric ones */
static double cones_step_length(oipm_t *S, const double *dz_, const double *ds_, const double *z_, const double *s_, double amax)
{
    double alpha = amax;
    for (idx k = 0; k < S->ncones; k++) {
        cone_t *c = &S->cones[k]; idx n = c->dim;
        if (cone_is_nonsym(c)) continue;
        const double *dz = dz_ + c->off, *ds = ds_ + c->off, *z = z_ + c->off, *s = s_ + c->off;
        double az = alpha, as = alpha;
        if (c->type == CONE_NONNEG) {
            for (idx i = 0; i < n; i++) {
                if (dz[i] < 0.0) { double t = -z[i] / dz[i]; if (t < az) az = t; }
                if (ds[i] < 0.0) { double t = -s[i] / ds[i]; if (t < as) as = t; }
            }
        } else if (c->type == CONE_SOC) {
            az = soc_step_component(z, dz, n, alpha);
            as = soc_step_component(s, ds, n, alpha);
        } else if (c->type == CONE_PSD) {
            psd_mul_Wx(c, 0, c->wv, dz, 1.0, 0.0, c->R);      az = psd_step_component(c, c->wv, alpha);
            psd_mul_Wx(c, 1, c->wv, ds, 1.0, 0.0, c->Rinv);   as = psd_step_component(c, c->wv, alpha);
        }
        double m = az < as ? az : as;
        if (m < alpha) alpha = m;
    }
    if (!S->all_symmetric) {
        double ceil_ = 1.0 - sqrt(NS_EPS);
        if (ceil_ < alpha) alpha = ceil_;
        for (idx k = 0; k < S->ncones; k++) {
            cone_t *c = &S->cones[k];
            if (!cone_is_nonsym(c)) continue;
            const double *dz = dz_ + c->off, *ds = ds_ + c->off, *z = z_ + c->off, *s = s_ + c->off;
            /* expcone.rs:154-174 */
            double az = ns3_backtrack_search(c, dz, z, alpha, S->set.min_terminate_step_length, S->set.linesearch_backtrack_step, 1);
            double as = ns3_backtrack_search(c, ds, s, alpha, S->set.min_terminate_step_length, S->set.linesearch_backtrack_step, 0);
            double m = az < as ? az : as;
            if (m < alpha) alpha = m;
        }
    }
    return alpha;
}

/* Cone::compute_barrier summed over the cones (compositecone.rs:334-345; nonnegativecone.rs:155-166,
   socone.rs:304-314, zerocone.rs:129-131, psdtrianglecone.rs:281-306, expcone.rs:176-187) */
static double cones_compute_barrier(oipm_t *S, const double *z_, const double *s_, const double *dz_, const double *ds_, double a)
{
    double barrier = 0.0;
    for (idx k = 0; k < S->ncones; k++) {
        cone_t *c = &S->cones[k]; idx n = c->dim;
        const double *dz = dz_ + c->off, *ds = ds_ + c->off, *z = z_ + c->off, *s = s_ + c->off;
        if (c->type == CONE_NONNEG) {
            double b = 0.0;
            for (idx i = 0; i < n; i++) { double si = s[i] + a * ds[i], zi = z[i] + a * dz[i]; b -= logsafe(si * zi); }
            barrier += b;
        } else if (c->type == CONE_SOC) {
            /* _soc_residual_shifted (socone.rs:410-417) */
            double rs, rz;
            { double x0 = s[0] + a * ds[0]; double *t = dvec(n); for (idx i = 1; i < n; i++) t[i - 1] = s[i] + a * ds[i]; double nm = vnorm(t, n - 1); free(t); rs = (x0 - nm) * (x0 + nm); }
            { double x0 = z[0] + a * dz[0]; double *t = dvec(n); for (idx i = 1; i < n; i++) t[i - 1] = z[i] + a * dz[i]; double nm = vnorm(t, n - 1); free(t); rz = (x0 - nm) * (x0 + nm); }
            barrier += (rs > 0.0 && rz > 0.0) ? -logsafe(rs * rz) * 0.5 : INFINITY;
        } else if (c->type == CONE_PSD) {
            double b = 0.0;
            for (int pass = 0; pass < 2; pass++) {
                const double *x = pass == 0 ? z : s, *dx = pass == 0 ? dz : ds;
                for (idx i = 0; i < n; i++) c->wv[i] = 1.0 * x[i] + a * dx[i];
                svec_to_mat(c->W1, c->psd_n, c->wv);
                if (chol_lower(c->W2, c->W1, c->psd_n) == 0) { double ld = 0.0; for (idx i = 0; i < c->psd_n; i++) ld += log(MAT(c->W2, c->psd_n, i, i)); b -= ld + ld; }
                else b -= INFINITY;
            }
            barrier += b;
        } else if (cone_is_ns3(c)) {
            double cz[3] = {z[0] + a * dz[0], z[1] + a * dz[1], z[2] + a * dz[2]};
            double cs[3] = {s[0] + a * ds[0], s[1] + a * ds[1], s[2] + a * ds[2]};
            double b = 0.0;
            if (c->type == CONE_EXP) { b += exp_barrier_dual(cz); b += exp_barrier_primal(cs); }
            else { b += pow_barrier_dual(cz, c->ns->alpha); b += pow_barrier_primal(cs, c->ns->alpha); }
            barrier += jit(b);
        } else if (c->type == CONE_GENPOW) {      /* genpowcone.rs:249-263: primal first, then dual */
            double b = 0.0, *w = c->gp->work;
            for (idx i = 0; i < n; i++) w[i] = 1.0 * s[i] + a * ds[i];
            b += gp_barrier_primal(c->gp, w);
            for (idx i = 0; i < n; i++) w[i] = 1.0 * z[i] + a * dz[i];
            b += gp_barrier_dual(c->gp, w);
            barrier += b;
        }
    }
    return barrier;
}

static void cones_margins(oipm_t *S, const double *z_, double *amin, double *bsum)
{
    double a = 1.7976931348623157e308, b = 0.0;
    for (idx k = 0; k < S->ncones; k++) {
        cone_t *c = &S->cones[k]; const double *z = z_ + c->off; idx n = c->dim;
        double ai = 1.7976931348623157e308, bi = 0.0;
        if (c->type == CONE_NONNEG) {
            ai = INFINITY;
            for (idx i = 0; i < n; i++) { if (z[i] < ai) ai = z[i]; bi += z[i] > 0.0 ? z[i] : 0.0; }
        } else if (c->type == CONE_SOC) {
            ai = z[0] - vnorm(z + 1, n - 1); bi = ai > 0.0 ? ai : 0.0;
        } else if (c->type == CONE_PSD && c->psd_n > 0) {
            double *ev = dvec(c->psd_n);
            svec_to_mat(c->W1, c->psd_n, z); sym_eigvals(c->W1, c->psd_n, ev);
            ai = INFINITY;
            for (idx i = 0; i < c->psd_n; i++) { if (ev[i] < ai) ai = ev[i]; bi += ev[i] > 0.0 ? ev[i] : 0.0; }
            free(ev);
        }
        if (ai < a) a = ai;
        b += bi;
    }
    *amin = a; *bsum = b;
}
static void cones_scaled_unit_shift(oipm_t *S, double *z_, double alpha, int primal)
{
    for (idx k = 0; k < S->ncones; k++) {
        cone_t *c = &S->cones[k]; double *z = z_ + c->off; idx n = c->dim;
        if (c->type == CONE_ZERO) { if (primal) for (idx i = 0; i < n; i++) z[i] = 0.0; }
        else if (c->type == CONE_NONNEG) for (idx i = 0; i < n; i++) z[i] += alpha;
        else if (c->type == CONE_PSD) for (idx k = 0; k < c->psd_n; k++) z[tri_index(k)] += alpha;
        else z[0] += alpha;
    }
}
static void shift_to_cone_interior(oipm_t *S, double *z, int primal)
{
    double minm, posm;
    cones_margins(S, z, &minm, &posm);
    double target = (posm * 0.1) / (double)S->degree;
    if (!(target > 1.0)) target = 1.0;   /* T::max(1, x): NaN-safe like f64::max */
    if (minm <= 0.0) { cones_scaled_unit_shift(S, z, -minm, primal); cones_scaled_unit_shift(S, z, target, primal); }
    else if (minm < target) cones_scaled_unit_shift(S, z, target - minm, primal);
    else cones_scaled_unit_shift(S, z, 0.0, primal);
}

/* ------------------------------------------------------------ KKT matrix */
static void kkt_assemble(oipm_t *S)
{
    idx n = S->n, m = S->m;
    const csc *P = &S->P, *A = &S->A;
    idx p = 0, nnz_vec = 0, nHs = 0;
    for (idx k = 0; k < S->ncones; k++) {
        cone_t *c = &S->cones[k];
        c->boff = nHs;
        c->blen = cone_Hs_diag(c) ? c->dim : c->dim * (c->dim + 1) / 2;
        nHs += c->blen;
        if (cone_pdim(c)) { p += cone_pdim(c); nnz_vec += 2 * c->dim; }   /* SOC: u, v; GenPow: q (dim1) + r (dim2) + p (dim) */
    }
    S->p = p; S->nHs = nHs;
    idx N = n + m + p; S->N = N;
    idx ndiagP = 0;
    for (idx i = 0; i < n; i++)
        if (P->colptr[i + 1] != P->colptr[i] && P->rowval[P->colptr[i + 1] - 1] == i) ndiagP++;
    idx nnzK = P->colptr[n] + n - ndiagP + A->colptr[n] + nHs + nnz_vec + p;
    csc *K = &S->K;
    K->m = K->n = N; K->colptr = ivec(N + 2); K->rowval = ivec(nnzK); K->nzval = dvec(nnzK);
    S->map_P = ivec(P->colptr[n]); S->map_A = ivec(A->colptr[n]); S->map_Hs = ivec(nHs);
    S->map_diagP = ivec(n); S->map_diag_full = ivec(N);
    idx *cp = K->colptr;
    /* column counts */
    for (idx i = 0; i < n; i++) cp[i] += P->colptr[i + 1] - P->colptr[i];
    for (idx i = 0; i < n; i++)
        if (P->colptr[i] == P->colptr[i + 1] || P->rowval[P->colptr[i + 1] - 1] != i) cp[i] += 1;
    for (idx q = 0; q < A->colptr[n]; q++) cp[n + A->rowval[q]] += 1;
    idx pcol = m + n;
    for (idx k = 0; k < S->ncones; k++) {
        cone_t *c = &S->cones[k]; idx row = c->off + n;
        if (cone_Hs_diag(c)) for (idx i = 0; i < c->dim; i++) cp[row + i] += 1;
        else for (idx i = 0; i < c->dim; i++) cp[row + i] += i + 1;
        if (cone_is_sparse(c)) { cp[pcol] += c->dim; cp[pcol + 1] += c->dim; cp[pcol] += 1; cp[pcol + 1] += 1; pcol += 2; }
        if (c->type == CONE_GENPOW) {   /* datamaps.rs:264-287: q, r, p columns + their diagonal entries */
            cp[pcol] += c->gp->dim1 + 1; cp[pcol + 1] += c->gp->dim2 + 1; cp[pcol + 2] += c->dim + 1; pcol += 3;
        }
    }
    /* counts -> pointers (next-fill positions) */
    { idx cur = 0; for (idx j = 0; j <= N; j++) { idx cnt = cp[j]; cp[j] = cur; cur += cnt; } }
    /* fill P (N shape), missing diagonal, A transposed */
    for (idx i = 0; i < n; i++)
        for (idx q = P->colptr[i]; q < P->colptr[i + 1]; q++) {
            idx dest = cp[i]++; K->rowval[dest] = P->rowval[q]; K->nzval[dest] = P->nzval[q]; S->map_P[q] = dest;
        }
    for (idx i = 0; i < n; i++)
        if (P->colptr[i] == P->colptr[i + 1] || P->rowval[P->colptr[i + 1] - 1] != i) {
            idx dest = cp[i]++; K->rowval[dest] = i; K->nzval[dest] = 0.0;
        }
    for (idx i = 0; i < A->n; i++)
        for (idx q = A->colptr[i]; q < A->colptr[i + 1]; q++) {
            idx col = A->rowval[q] + n; idx dest = cp[col]++;
            K->rowval[dest] = i; K->nzval[dest] = A->nzval[q]; S->map_A[q] = dest;
        }
    pcol = m + n;
    for (idx k = 0; k < S->ncones; k++) {
        cone_t *c = &S->cones[k]; idx row = c->off + n; idx *blk = S->map_Hs + c->boff;
        if (cone_Hs_diag(c)) {
            for (idx i = 0; i < c->dim; i++) { idx col = row + i; idx dest = cp[col]++; K->rowval[dest] = col; K->nzval[dest] = 0.0; blk[i] = dest; }
        } else {
            idx kidx = 0;
            for (idx col = row; col < row + c->dim; col++)
                for (idx r = row; r <= col; r++) { idx dest = cp[col]++; K->rowval[dest] = r; K->nzval[dest] = 0.0; blk[kidx++] = dest; }
        }
        if (cone_is_sparse(c)) {
            c->map_u = ivec(c->dim); c->map_v = ivec(c->dim);
            /* v is the first extra column, u the second (datamaps.rs:186-189) */
            for (idx i = 0; i < c->dim; i++) { idx dest = cp[pcol]++; K->rowval[dest] = row + i; K->nzval[dest] = 0.0; c->map_v[i] = dest; }
            for (idx i = 0; i < c->dim; i++) { idx dest = cp[pcol + 1]++; K->rowval[dest] = row + i; K->nzval[dest] = 0.0; c->map_u[i] = dest; }
            for (idx i = 0; i < 2; i++) { idx col = pcol + i; idx dest = cp[col]++; K->rowval[dest] = col; K->nzval[dest] = 0.0; c->map_D[i] = dest; }
            pcol += 2;
        }
        if (c->type == CONE_GENPOW) {   /* datamaps.rs:289-312 */
            gp_t *g = c->gp;
            g->map_q = ivec(g->dim1); g->map_r = ivec(g->dim2); g->map_p = ivec(c->dim);
            for (idx i = 0; i < g->dim1; i++) { idx dest = cp[pcol]++; K->rowval[dest] = row + i; K->nzval[dest] = 0.0; g->map_q[i] = dest; }
            for (idx i = 0; i < g->dim2; i++) { idx dest = cp[pcol + 1]++; K->rowval[dest] = row + g->dim1 + i; K->nzval[dest] = 0.0; g->map_r[i] = dest; }
            for (idx i = 0; i < c->dim; i++) { idx dest = cp[pcol + 2]++; K->rowval[dest] = row + i; K->nzval[dest] = 0.0; g->map_p[i] = dest; }
            for (idx i = 0; i < 3; i++) { idx col = pcol + i; idx dest = cp[col]++; K->rowval[dest] = col; K->nzval[dest] = 0.0; g->map_D[i] = dest; }
            pcol += 3;
        }
    }
    /* backshift */
    for (idx j = N; j > 0; j--) cp[j] = cp[j - 1];
    cp[0] = 0;
    for (idx j = 0; j < N; j++) S->map_diag_full[j] = cp[j + 1] - 1;
    for (idx j = 0; j < n; j++) S->map_diagP[j] = cp[j + 1] - 1;
    /* signs */
    S->dsigns = (int8_t *)malloc((size_t)N);
    for (idx i = 0; i < N; i++) S->dsigns[i] = 1;
    for (idx i = n; i < n + m; i++) S->dsigns[i] = -1;
    idx pp = n + m;
    for (idx k = 0; k < S->ncones; k++) {
        if (cone_is_sparse(&S->cones[k])) { S->dsigns[pp] = -1; S->dsigns[pp + 1] = 1; pp += 2; }
        else if (S->cones[k].type == CONE_GENPOW) { S->dsigns[pp] = -1; S->dsigns[pp + 1] = -1; S->dsigns[pp + 2] = 1; pp += 3; }   /* datamaps.rs:252-254 */
    }
    S->info.nnzK = nnzK;
}

static void kkt_update_values(oipm_t *S, const idx *index, const double *v, idx len)
{ for (idx i = 0; i < len; i++) S->K.nzval[index[i]] = v[i]; oq_update_values(S->ldl, index, v, len); }
static void kkt_scale_values(oipm_t *S, const idx *index, idx len, double sc)
{ for (idx i = 0; i < len; i++) S->K.nzval[index[i]] *= sc; oq_scale_values(S->ldl, index, len, sc); }

static int kkt_regularize_and_refactor(oipm_t *S)
{
    double *diag_kkt = S->kw1, *diag_shift = S->kw2; idx N = S->N;
    if (S->set.static_regularization_enable) {
        for (idx i = 0; i < N; i++) diag_kkt[i] = S->K.nzval[S->map_diag_full[i]];
        double eps = S->set.static_regularization_constant + S->set.static_regularization_proportional * vnorm_inf(diag_kkt, N);
        for (idx i = 0; i < N; i++) diag_shift[i] = diag_kkt[i];
        for (idx i = 0; i < N; i++) { if (S->dsigns[i] == 1) diag_shift[i] += eps; else diag_shift[i] -= eps; }
        kkt_update_values(S, S->map_diag_full, diag_shift, N);
        S->diagonal_regularizer = eps;
    }
    int rc = oq_refactor(S->ldl);
    S->info.n_refactor++;
    int ok = (rc == 0) && oq_dinv_is_finite(S->ldl);
    if (!ok && getenv("OIPM_DEBUG")) fprintf(stderr, "[oipm] refactor failed rc=%d finite=%d regcount=%lld\n", rc, oq_dinv_is_finite(S->ldl), (long long)oq_regularize_count(S->ldl));
    if (S->set.static_regularization_enable)
        for (idx i = 0; i < N; i++) S->K.nzval[S->map_diag_full[i]] = diag_kkt[i];
    return ok;
}

static int kkt_update(oipm_t *S)
{
    cones_get_Hs(S, S->Hs);
    for (idx i = 0; i < S->nHs; i++) S->Hs[i] = -S->Hs[i];
    kkt_update_values(S, S->map_Hs, S->Hs, S->nHs);
    for (idx k = 0; k < S->ncones; k++) {
        cone_t *c = &S->cones[k];
        if (c->type == CONE_GENPOW) {   /* datamaps.rs:314-337: sqrt(mu) distributed to the off-diagonal vectors */
            gp_t *g = c->gp; double sq = sqrt(g->mu);
            kkt_update_values(S, g->map_q, g->q, g->dim1);
            kkt_update_values(S, g->map_r, g->r, g->dim2);
            kkt_update_values(S, g->map_p, g->p, c->dim);
            kkt_scale_values(S, g->map_q, g->dim1, -sq);
            kkt_scale_values(S, g->map_r, g->dim2, -sq);
            kkt_scale_values(S, g->map_p, c->dim, -sq);
            double d3[3] = {-1.0, -1.0, 1.0};
            kkt_update_values(S, g->map_D, d3, 3);
            continue;
        }
        if (!cone_is_sparse(c)) continue;
        double e2 = c->eta * c->eta;
        kkt_update_values(S, c->map_u, c->u, c->dim);
        kkt_update_values(S, c->map_v, c->v, c->dim);
        kkt_scale_values(S, c->map_u, c->dim, -e2);
        kkt_scale_values(S, c->map_v, c->dim, -e2);
        double dd[2] = {-e2, e2};
        kkt_update_values(S, c->map_D, dd, 2);
    }
    return kkt_regularize_and_refactor(S);
}

static void kkt_setrhs(oipm_t *S, const double *rx, const double *rz)
{
    memcpy(S->kb, rx, (size_t)S->n * sizeof(double));
    memcpy(S->kb + S->n, rz, (size_t)S->m * sizeof(double));
    for (idx i = S->n + S->m; i < S->N; i++) S->kb[i] = 0.0;
}

static void ldl_solve(oipm_t *S, double *x, const double *b)
{ memcpy(x, b, (size_t)S->N * sizeof(double)); oq_solve(S->ldl, x); S->info.n_ldl_solve++; }

static double refine_error(oipm_t *S, double *e, const double *b, const double *xi)
{ memcpy(e, b, (size_t)S->N * sizeof(double)); symv_tri(&S->K, e, xi, -1.0, 1.0); return vnorm_inf(e, S->N); }

static int kkt_iterative_refinement(oipm_t *S)
{
    double *x = S->kx, *b = S->kb, *e = S->kw1, *dx = S->kw2; idx N = S->N;
    double normb = vnorm_inf(b, N);
    double norme = refine_error(S, e, b, x);
    if (!isfinite(norme)) { if (getenv("OIPM_DEBUG")) fprintf(stderr, "[oipm] IR: initial residual not finite (normb %g)\n", normb); return 0; }
    for (int it = 0; it < S->set.iterative_refinement_max_iter; it++) {
        if (norme <= S->set.iterative_refinement_abstol + S->set.iterative_refinement_reltol * normb) break;
        double last = norme;
        ldl_solve(S, dx, e);
        for (idx i = 0; i < N; i++) dx[i] = 1.0 * x[i] + 1.0 * dx[i];
        norme = refine_error(S, e, b, dx);
        if (!isfinite(norme)) return 0;
        double ratio = last / norme;
        if (ratio < S->set.iterative_refinement_stop_ratio) {
            if (ratio > 1.0) { double *t = x; x = dx; dx = t; }
            break;
        }
        { double *t = x; x = dx; dx = t; }
    }
    if (x != S->kx) { S->kw2 = S->kx; S->kx = x; }  /* std::mem::swap of the buffers */
    return 1;
}

static int kkt_solve(oipm_t *S, double *lhsx, double *lhsz)
{
    ldl_solve(S, S->kx, S->kb);
    int ok = S->set.iterative_refinement_enable ? kkt_iterative_refinement(S) : vfinite(S->kx, S->N);
    if (ok) {
        if (lhsx) memcpy(lhsx, S->kx, (size_t)S->n * sizeof(double));
        if (lhsz) memcpy(lhsz, S->kx + S->n, (size_t)S->m * sizeof(double));
    }
    return ok;
}

/* --------------------------------------------------------- equilibration */
static void scale_data(oipm_t *S, const double *d, const double *e)
{
    csc *P = &S->P, *A = &S->A;
    if (d) {
        for (idx col = 0; col < P->n; col++) for (idx p = P->colptr[col]; p < P->colptr[col + 1]; p++) P->nzval[p] *= d[P->rowval[p]] * d[col];
        for (idx col = 0; col < A->n; col++) for (idx p = A->colptr[col]; p < A->colptr[col + 1]; p++) A->nzval[p] *= e[A->rowval[p]] * d[col];
        for (idx i = 0; i < S->n; i++) S->q[i] *= d[i];
    } else {
        for (idx p = 0; p < A->colptr[A->n]; p++) A->nzval[p] *= e[A->rowval[p]];
    }
    for (idx i = 0; i < S->m; i++) S->b[i] *= e[i];
}

static void equilibrate(oipm_t *S)
{
    idx n = S->n, m = S->m;
    if (!S->set.equilibrate_enable) return;
    double *d = S->d, *e = S->e, *dw = S->dinv, *ew = S->einv;
    csc *P = &S->P, *A = &S->A;
    double smin = S->set.equilibrate_min_scaling, smax = S->set.equilibrate_max_scaling;
    for (int it = 0; it < S->set.equilibrate_max_iter; it++) {
        for (idx i = 0; i < n; i++) dw[i] = 0.0;
        for (idx i = 0; i < n; i++) for (idx p = P->colptr[i]; p < P->colptr[i + 1]; p++) {
            double t = fabs(P->nzval[p]); idx r = P->rowval[p];
            if (t > dw[i]) dw[i] = t;
            if (t > dw[r]) dw[r] = t;
        }
        for (idx i = 0; i < n; i++) for (idx p = A->colptr[i]; p < A->colptr[i + 1]; p++) { double t = fabs(A->nzval[p]); if (t > dw[i]) dw[i] = t; }
        for (idx i = 0; i < m; i++) ew[i] = 0.0;
        for (idx p = 0; p < A->colptr[n]; p++) { double t = fabs(A->nzval[p]); idx r = A->rowval[p]; if (t > ew[r]) ew[r] = t; }
        for (idx i = 0; i < n; i++) if (dw[i] == 0.0) dw[i] = 1.0;
        for (idx i = 0; i < m; i++) if (ew[i] == 0.0) ew[i] = 1.0;
        for (idx i = 0; i < n; i++) dw[i] = 1.0 / sqrt(dw[i]);
        for (idx i = 0; i < m; i++) ew[i] = 1.0 / sqrt(ew[i]);
        for (idx i = 0; i < n; i++) dw[i] = clipd(dw[i], smin / d[i], smax / d[i]);
        for (idx i = 0; i < m; i++) ew[i] = clipd(ew[i], smin / e[i], smax / e[i]);
        scale_data(S, dw, ew);
        for (idx i = 0; i < n; i++) d[i] *= dw[i];
        for (idx i = 0; i < m; i++) e[i] *= ew[i];
        /* cost scaling: plain (non-symmetric) column norms of the triu P */
        for (idx i = 0; i < n; i++) { double v = 0.0; for (idx p = P->colptr[i]; p < P->colptr[i + 1]; p++) { double t = fabs(P->nzval[p]); if (t > v) v = t; } dw[i] = v; }
        double meanP = vmean(dw, n), infq = vnorm_inf(S->q, n);
        if (meanP != 0.0 && infq != 0.0) {
            double sc = infq > meanP ? infq : meanP;
            double ct = clipd(1.0 / sc, smin / S->c, smax / S->c);
            for (idx p = 0; p < P->colptr[n]; p++) P->nzval[p] *= ct;
            for (idx i = 0; i < n; i++) S->q[i] *= ct;
            S->c *= ct;
        }
    }
    /* rectification: SOC cones need a scalar scaling (socone.rs:97-101) */
    int changed = 0;
    for (idx i = 0; i < m; i++) ew[i] = 1.0;
    for (idx k = 0; k < S->ncones; k++) {
        cone_t *c = &S->cones[k];
        if (c->type == CONE_SOC || c->type == CONE_PSD || cone_is_nonsym(c)) {
            double mean = vmean(e + c->off, c->dim);
            for (idx i = 0; i < c->dim; i++) ew[c->off + i] = (1.0 / e[c->off + i]) * mean;
            changed = 1;
        }
    }
    if (changed) { scale_data(S, NULL, ew); for (idx i = 0; i < m; i++) e[i] *= ew[i]; }
    for (idx i = 0; i < n; i++) S->dinv[i] = 1.0 / d[i];
    for (idx i = 0; i < m; i++) S->einv[i] = 1.0 / e[i];
}

/* -------------------------------------------------------------- lifecycle */
void oipm_free(oipm_t *S)
{
    if (!S) return;
    csc_free(&S->P); csc_free(&S->A); free(S->q); free(S->b); free(S->keep);
    for (idx k = 0; k < S->ncones; k++) { cone_t *c = &S->cones[k]; free(c->w); free(c->lam); free(c->u); free(c->v); free(c->map_u); free(c->map_v);
        free(c->R); free(c->Rinv); free(c->lisqrt); free(c->HsM); free(c->W1); free(c->W2); free(c->W3); free(c->wv); free(c->ns);
        if (c->gp) { gp_t *g = c->gp; free(g->alpha); free(g->grad); free(g->z); free(g->p); free(g->q); free(g->r); free(g->d1); free(g->work); free(g->work_pb); free(g->map_p); free(g->map_q); free(g->map_r); free(g); } }
    free(S->cones); free(S->d); free(S->dinv); free(S->e); free(S->einv);
    if (S->K.colptr) csc_free(&S->K);
    free(S->map_P); free(S->map_A); free(S->map_Hs); free(S->map_diagP); free(S->map_diag_full);
    free(S->dsigns); free(S->Hs); free(S->kx); free(S->kb); free(S->kw1); free(S->kw2);
    if (S->ldl) oq_free(S->ldl);
    free(S->perm);
    free(S->x1); free(S->z1); free(S->x2); free(S->z2); free(S->workx); free(S->workz); free(S->work_conic);
    free(S->vx); free(S->vs); free(S->vz); free(S->lx); free(S->ls); free(S->lz);
    free(S->rx_); free(S->rs_); free(S->rz_); free(S->px); free(S->ps); free(S->pz);
    free(S->rx); free(S->rz); free(S->rx_inf); free(S->rz_inf); free(S->Px);
    free(S);
}

/* P must be upper triangular CSC (the reference converts with to_triu,
   problemdata.rs:79-81; the Python caller does the same). */
int oipm_new_ex(oipm_t **out, idx n, idx m, const idx *Pp, const idx *Pi, const double *Px,
                const double *q, const idx *Ap, const idx *Ai, const double *Ax, const double *b,
                idx ncones_in, const int32_t *ctype, const idx *cdim, const double *cparam, const oipm_settings *set);
int oipm_new(oipm_t **out, idx n, idx m, const idx *Pp, const idx *Pi, const double *Px,
             const double *q, const idx *Ap, const idx *Ai, const double *Ax, const double *b,
             idx ncones_in, const int32_t *ctype, const idx *cdim, const oipm_settings *set)
{
    return oipm_new_ex(out, n, m, Pp, Pi, Px, q, Ap, Ai, Ax, b, ncones_in, ctype, cdim, NULL, set);
}
int oipm_new_gp(oipm_t **out, idx n, idx m, const idx *Pp, const idx *Pi, const double *Px,
                const double *q, const idx *Ap, const idx *Ai, const double *Ax, const double *b,
                idx ncones_in, const int32_t *ctype, const idx *cdim, const double *cparam,
                const idx *gp_dim2, const double *gp_alpha, const oipm_settings *set);
/* cparam[k]: the exponent of a PowerConeT(alpha) (supportedcone.rs:36-38), ignored for the other cones */
int oipm_new_ex(oipm_t **out, idx n, idx m, const idx *Pp, const idx *Pi, const double *Px,
                const double *q, const idx *Ap, const idx *Ai, const double *Ax, const double *b,
                idx ncones_in, const int32_t *ctype, const idx *cdim, const double *cparam, const oipm_settings *set)
{
    return oipm_new_gp(out, n, m, Pp, Pi, Px, q, Ap, Ai, Ax, b, ncones_in, ctype, cdim, cparam, NULL, NULL, set);
}
/* GenPowerConeT(alpha, dim2) (supportedcone.rs:44): cdim[k] = len(alpha), gp_dim2[k] = dim2, the exponents of all
   such cones concatenated in cone order in gp_alpha */
int oipm_new_gp(oipm_t **out, idx n, idx m, const idx *Pp, const idx *Pi, const double *Px,
                const double *q, const idx *Ap, const idx *Ai, const double *Ax, const double *b,
                idx ncones_in, const int32_t *ctype, const idx *cdim, const double *cparam,
                const idx *gp_dim2, const double *gp_alpha, const oipm_settings *set)
{
    *out = NULL;
    idx gp_cursor = 0;
    oipm_t *S = (oipm_t *)calloc(1, sizeof(oipm_t));
    if (set) S->set = *set; else oipm_default_settings(&S->set);
    S->n = n; S->m = m;
    csc_copy(&S->P, n, n, Pp, Pi, Px);
    csc_copy(&S->A, m, n, Ap, Ai, Ax);
    S->q = dvec(n); memcpy(S->q, q, (size_t)n * sizeof(double));
    S->b = dvec(m); memcpy(S->b, b, (size_t)m * sizeof(double));
    for (idx i = 0; i < m; i++) if (S->b[i] > g_infinity) S->b[i] = g_infinity;
    /* collapse cones */
    S->cones = (cone_t *)calloc((size_t)(ncones_in > 0 ? ncones_in : 1), sizeof(cone_t));
    idx nc = 0, k = 0;
    while (k < ncones_in) {
        int t = ctype[k]; idx dm = cdim[k];
        if (t == CONE_EXP || t == CONE_POW) dm = 3;
        idx gdim1 = 0, gdim2 = 0; const double *galpha = NULL;
        if (t == CONE_GENPOW) {
            if (!gp_dim2 || !gp_alpha) { oipm_free(S); return -3; }
            gdim1 = cdim[k]; gdim2 = gp_dim2[k]; galpha = gp_alpha + gp_cursor; gp_cursor += gdim1;
            dm = gdim1 + gdim2;
        }
        idx numel = (t == CONE_PSD) ? dm * (dm + 1) / 2 : dm;
        if (numel == 0) { k++; continue; }
        int collapsible = (t == CONE_NONNEG) || (t == CONE_SOC && dm == 1) || (t == CONE_PSD && dm == 1);
        if (collapsible) {
            idx tot = (t == CONE_NONNEG) ? dm : 1;
            k++;
            while (k < ncones_in) {
                int t2 = ctype[k]; idx d2 = cdim[k];
                idx ne2 = (t2 == CONE_PSD) ? d2 * (d2 + 1) / 2 : d2;
                if (ne2 != 0) {
                    if (t2 == CONE_NONNEG) tot += d2;
                    else if ((t2 == CONE_SOC || t2 == CONE_PSD) && d2 == 1) tot += 1;
                    else break;
                }
                k++;
            }
            S->cones[nc].type = CONE_NONNEG; S->cones[nc].dim = tot; nc++;
        } else {
            if (t == CONE_SOC && dm < 2) { oipm_free(S); return -2; }
            S->cones[nc].type = t; S->cones[nc].dim = (t == CONE_PSD) ? dm * (dm + 1) / 2 : dm;
            S->cones[nc].psd_n = (t == CONE_PSD) ? dm : 0;
            if (t == CONE_EXP || t == CONE_POW) {
                S->cones[nc].ns = (ns3_t *)calloc(1, sizeof(ns3_t));
                S->cones[nc].ns->alpha = (t == CONE_POW && cparam) ? cparam[k] : 0.5;
                if (t == CONE_POW && !(S->cones[nc].ns->alpha > 0.0 && S->cones[nc].ns->alpha < 1.0)) { oipm_free(S); return -3; }
            }
            if (t == CONE_GENPOW) {       /* GenPowerConeData::new (genpowcone.rs:41-60) */
                gp_t *g = (gp_t *)calloc(1, sizeof(gp_t));
                S->cones[nc].gp = g;
                g->dim1 = gdim1; g->dim2 = gdim2;
                g->alpha = dvec(gdim1); memcpy(g->alpha, galpha, (size_t)gdim1 * sizeof(double));
                double asum = 0.0, asq = 0.0; int pos = 1;
                for (idx i = 0; i < gdim1; i++) { asum += g->alpha[i]; asq += g->alpha[i] * g->alpha[i]; if (!(g->alpha[i] > 0.0)) pos = 0; }
                if (!pos || gdim1 < 1 || !(fabs(1.0 - asum) < 2.220446049250313e-16 * (double)gdim1 * 0.5 + 1e-300)) { oipm_free(S); return -3; }
                g->grad = dvec(dm); g->z = dvec(dm); g->p = dvec(dm); g->q = dvec(gdim1); g->r = dvec(gdim2); g->d1 = dvec(gdim1);
                g->mu = 1.0; g->d2 = 0.0; g->psi = 1.0 / asq; g->work = dvec(dm); g->work_pb = dvec(dm);
            }
            nc++; k++;
        }
    }
    S->ncones = nc;
    /* presolve (presolver.rs:157-204, 75-125; problemdata.rs:86-93): nonnegative rows whose bound is infinite are
       dropped from A, b and their cone.  b was capped at the bound above, which still compares as "beyond" it. */
    S->mfull = m; S->keep = NULL;
    if (S->set.presolve_enable) {
        const double thr = (1.0 - 2.220446049250313e-16 * 10.0) * g_infinity;
        char *keep = (char *)malloc((size_t)(m > 0 ? m : 1));
        idx mred = m, r = 0;
        for (idx i = 0; i < m; i++) keep[i] = 1;
        for (idx c = 0; c < nc; c++) {
            cone_t *cn = &S->cones[c];
            if (cn->type == CONE_NONNEG) { for (idx i = 0; i < cn->dim; i++, r++) if (S->b[r] > thr) { keep[r] = 0; mred--; } }
            else r += cn->dim;
        }
        if (r != m) { free(keep); oipm_free(S); return -1; }
        if (mred < m) {
            idx nc2 = 0; r = 0;
            for (idx c = 0; c < nc; c++) {
                cone_t cn = S->cones[c];
                if (cn.type == CONE_NONNEG) {
                    idx nk = 0;
                    for (idx i = 0; i < cn.dim; i++) nk += keep[r + i];
                    r += cn.dim;
                    if (nk > 0) { cn.dim = nk; S->cones[nc2++] = cn; }
                } else { r += cn.dim; S->cones[nc2++] = cn; }
            }
            for (idx c = nc2; c < nc; c++) memset(&S->cones[c], 0, sizeof(cone_t));
            nc = nc2; S->ncones = nc;
            idx *rowmap = ivec(m), nr = 0;
            for (idx i = 0; i < m; i++) rowmap[i] = keep[i] ? nr++ : -1;
            csc *A = &S->A; idx w = 0;
            for (idx j = 0; j < n; j++) {
                idx b0 = A->colptr[j]; A->colptr[j] = w;
                for (idx q_ = b0; q_ < A->colptr[j + 1]; q_++)
                    if (rowmap[A->rowval[q_]] >= 0) { A->rowval[w] = rowmap[A->rowval[q_]]; A->nzval[w] = A->nzval[q_]; w++; }
            }
            A->colptr[n] = w; A->m = mred;
            for (idx i = 0; i < m; i++) if (keep[i]) S->b[rowmap[i]] = S->b[i];
            free(rowmap);
            S->keep = keep; m = mred; S->m = mred;
        } else free(keep);
    }
    idx off = 0; S->degree = 0; S->all_symmetric = 1; S->allows_primal_dual = 1;
    for (idx c = 0; c < nc; c++) {
        cone_t *cn = &S->cones[c];
        if (cone_is_nonsym(cn)) S->all_symmetric = 0;
        if (cn->type == CONE_GENPOW) S->allows_primal_dual = 0;   /* genpowcone.rs:108-110 */
        cn->off = off; off += cn->dim;
        S->degree += cone_degree(cn);
        if (cn->type != CONE_ZERO && !cone_is_nonsym(cn)) { cn->w = dvec(cn->dim); cn->lam = dvec(cn->dim); }
        if (cn->type == CONE_PSD) {
            idx n2 = cn->psd_n * cn->psd_n;
            cn->R = dvec(n2); cn->Rinv = dvec(n2); cn->lisqrt = dvec(cn->psd_n); cn->HsM = dvec(cn->dim * cn->dim);
            cn->W1 = dvec(n2); cn->W2 = dvec(n2); cn->W3 = dvec(n2); cn->wv = dvec(cn->dim);
        }
        if (cn->type == CONE_SOC && cn->dim > 4) { cn->sparse = 1; cn->u = dvec(cn->dim); cn->v = dvec(cn->dim); cn->d = 1.0; }
    }
    if (off != m) { oipm_free(S); return -1; }
    S->normq = vnorm_inf(S->q, n); S->normb = vnorm_inf(S->b, m);
    S->d = dvec(n); S->dinv = dvec(n); S->e = dvec(m); S->einv = dvec(m); S->c = 1.0;
    for (idx i = 0; i < n; i++) { S->d[i] = 1.0; S->dinv[i] = 1.0; }
    for (idx i = 0; i < m; i++) { S->e[i] = 1.0; S->einv[i] = 1.0; }
    equilibrate(S);
    kkt_assemble(S);
    idx N = S->N;
    S->Hs = dvec(S->nHs); S->kx = dvec(N); S->kb = dvec(N); S->kw1 = dvec(N); S->kw2 = dvec(N);
    S->x1 = dvec(n); S->z1 = dvec(m); S->x2 = dvec(n); S->z2 = dvec(m);
    S->workx = dvec(n); S->workz = dvec(m); S->work_conic = dvec(m);
    S->vx = dvec(n); S->vs = dvec(m); S->vz = dvec(m); S->lx = dvec(n); S->ls = dvec(m); S->lz = dvec(m);
    S->rx_ = dvec(n); S->rs_ = dvec(m); S->rz_ = dvec(m); S->px = dvec(n); S->ps = dvec(m); S->pz = dvec(m);
    S->rx = dvec(n); S->rz = dvec(m); S->rx_inf = dvec(n); S->rz_inf = dvec(m); S->Px = dvec(n);
    S->vtau = S->vkap = 1.0; S->rtau = 1.0;
    *out = S;
    return 0;
}

/* KKT pattern for the ordering step (caller computes a permutation of size N) */
idx oipm_kkt_dim(const oipm_t *S) { return S->N; }
/* dynamically regularised pivots of the LAST refactorisation (qdldl.rs:104-112 regularize_count) */
idx oipm_regularize_count(const oipm_t *S) { return S->ldl ? oq_regularize_count(S->ldl) : 0; }
idx oipm_m_reduced(const oipm_t *S) { return S->m; }
idx oipm_kkt_nnz(const oipm_t *S) { return S->K.colptr[S->N]; }
const idx *oipm_kkt_colptr(const oipm_t *S) { return S->K.colptr; }
const idx *oipm_kkt_rowval(const oipm_t *S) { return S->K.rowval; }
const double *oipm_kkt_nzval(const oipm_t *S) { return S->K.nzval; }
const int8_t *oipm_kkt_dsigns(const oipm_t *S) { return S->dsigns; }
const idx *oipm_map(const oipm_t *S, int which, idx *len)
{
    switch (which) {
        case 0: *len = S->P.colptr[S->n]; return S->map_P;
        case 1: *len = S->A.colptr[S->n]; return S->map_A;
        case 2: *len = S->nHs; return S->map_Hs;
        case 3: *len = S->n; return S->map_diagP;
        case 4: *len = S->N; return S->map_diag_full;
    }
    *len = 0; return NULL;
}
/* sparse-cone maps: which = 0 u, 1 v, 2 D for the k-th sparse cone */
const idx *oipm_sparse_map(const oipm_t *S, idx ksparse, int which, idx *len)
{
    idx cnt = 0;
    for (idx k = 0; k < S->ncones; k++) {
        const cone_t *c = &S->cones[k];
        if (!cone_is_sparse(c)) continue;
        if (cnt == ksparse) {
            if (which == 0) { *len = c->dim; return c->map_u; }
            if (which == 1) { *len = c->dim; return c->map_v; }
            *len = 2; return c->map_D;
        }
        cnt++;
    }
    *len = 0; return NULL;
}
/* maps of the k-th generalised power cone: which 0 q, 1 r, 2 p, 3 D (datamaps.rs:226-243) */
const idx *oipm_genpow_map(const oipm_t *S, idx kgp, int which, idx *len)
{
    idx cnt = 0;
    for (idx k = 0; k < S->ncones; k++) {
        const cone_t *c = &S->cones[k];
        if (c->type != CONE_GENPOW) continue;
        if (cnt == kgp) {
            if (which == 0) { *len = c->gp->dim1; return c->gp->map_q; }
            if (which == 1) { *len = c->gp->dim2; return c->gp->map_r; }
            if (which == 2) { *len = c->dim; return c->gp->map_p; }
            *len = 3; return c->gp->map_D;
        }
        cnt++;
    }
    *len = 0; return NULL;
}
/* KKTSolver::update on the current cone scalings (needs set_perm); K values are then readable through oipm_kkt_nzval */
int oipm_test_kkt_update(oipm_t *S) { return S->ldl ? kkt_update(S) : -1; }
const double *oipm_equil(const oipm_t *S, int which) { return which == 0 ? S->d : (which == 1 ? S->e : &S->c); }
const double *oipm_scaled_data(const oipm_t *S, int which)
{ return which == 0 ? S->P.nzval : which == 1 ? S->A.nzval : which == 2 ? S->q : S->b; }

int oipm_set_perm(oipm_t *S, const idx *perm)
{
    if (S->ldl) { oq_free(S->ldl); S->ldl = NULL; }
    free(S->perm);
    S->perm = ivec(S->N);
    memcpy(S->perm, perm, (size_t)S->N * sizeof(idx));
    int rc = oq_new(&S->ldl, S->N, S->N, S->K.colptr, S->K.rowval, S->K.nzval, S->perm, S->dsigns, 1, 1,
                    S->set.dynamic_regularization_eps, S->set.dynamic_regularization_delta);
    if (rc) return rc;
    S->info.nnzL = oq_nnzL(S->ldl);
    return 0;
}

/* ------------------------------------------------------------ IPM pieces */
static void residuals_update(oipm_t *S)
{
    idx n = S->n, m = S->m;
    double qx = vdot(S->q, S->vx, n), bz = vdot(S->b, S->vz, m), sz = vdot(S->vs, S->vz, m);
    symv_tri(&S->P, S->Px, S->vx, 1.0, 0.0);
    double xPx = vdot(S->vx, S->Px, n);
    gemv_T(&S->A, S->rx_inf, S->vz, -1.0, 0.0);
    memcpy(S->rz_inf, S->vs, (size_t)m * sizeof(double));
    gemv_N(&S->A, S->rz_inf, S->vx, 1.0, 1.0);
    for (idx i = 0; i < n; i++) S->rx[i] = -1.0 * S->Px[i] + (-S->vtau) * S->q[i];
    for (idx i = 0; i < n; i++) S->rx[i] = 1.0 * S->rx_inf[i] + 1.0 * S->rx[i];
    for (idx i = 0; i < m; i++) S->rz[i] = 1.0 * S->rz_inf[i] + (-S->vtau) * S->b[i];
    S->rtau = qx + bz + S->vkap + xPx / S->vtau;
    S->dot_qx = qx; S->dot_bz = bz; S->dot_sz = sz; S->dot_xPx = xPx;
}

static void info_update(oipm_t *S, double t0)
{
    idx n = S->n, m = S->m; oipm_info *I = &S->info;
    double tinv = 1.0 / S->vtau, cinv = 1.0 / S->c;
    double xPx2 = S->dot_xPx * tinv * tinv / 2.0;
    I->cost_primal = (S->dot_qx * tinv + xPx2) * cinv;
    I->cost_dual = (-S->dot_bz * tinv - xPx2) * cinv;
    double normx = vnorm_scaled(S->vx, S->d, n), normz = vnorm_scaled(S->vz, S->e, m) * cinv, norms = vnorm_scaled(S->vs, S->einv, m);
    I->res_primal_inf = (vnorm_scaled(S->rx_inf, S->dinv, n) * cinv) / fmax(1.0, normz);
    I->res_dual_inf = fmax(vnorm_scaled(S->Px, S->dinv, n) / fmax(1.0, normx),
                           vnorm_scaled(S->rz_inf, S->einv, m) / fmax(1.0, normx + norms));
    normx *= tinv; normz *= tinv; norms *= tinv;
    I->res_primal = vnorm_scaled(S->rz, S->einv, m) * tinv / fmax(1.0, S->normb + normx + norms);
    I->res_dual = vnorm_scaled(S->rx, S->dinv, n) * tinv * cinv / fmax(1.0, S->normq + normx + normz);
    I->gap_abs = fabs(I->cost_primal - I->cost_dual);
    I->gap_rel = I->gap_abs / fmax(1.0, fmin(fabs(I->cost_primal), fabs(I->cost_dual)));
    I->ktratio = S->vkap * tinv;
    I->solve_time = now_s() - t0;
}

static void check_convergence(oipm_t *S, double tga, double tgr, double tf, double tia, double tir, double tkt,
                              int st_solved, int st_pinf, int st_dinf)
{
    oipm_info *I = &S->info;
    int solved = ((I->gap_abs < tga) || (I->gap_rel < tgr)) && (I->res_primal < tf) && (I->res_dual < tf);
    if (I->ktratio <= 1.0 && solved) I->status = st_solved;
    else if (I->ktratio > (1.0 / tkt) * 1000.0) {
        if ((S->dot_bz < -tia) && (I->res_primal_inf < -tir * S->dot_bz)) I->status = st_pinf;
        else if ((S->dot_qx < -tia) && (I->res_dual_inf < -tir * S->dot_qx)) I->status = st_dinf;
    }
}

static int check_termination(oipm_t *S, int iter)
{
    oipm_info *I = &S->info; const oipm_settings *T = &S->set;
    check_convergence(S, T->tol_gap_abs, T->tol_gap_rel, T->tol_feas, T->tol_infeas_abs, T->tol_infeas_rel,
                      T->tol_ktratio, ST_SOLVED, ST_PRIMAL_INFEASIBLE, ST_DUAL_INFEASIBLE);
    if (I->status == ST_UNSOLVED && iter > 1 &&
        (I->res_dual > S->prev_res_dual || I->res_primal > S->prev_res_primal)) {
        if (I->ktratio < 2.220446049250313e-16 * 100.0 &&
            (S->prev_gap_abs < T->tol_gap_abs || S->prev_gap_rel < T->tol_gap_rel))
            I->status = ST_INSUFFICIENT_PROGRESS;
        if (I->ktratio < 1.0) {
            if ((I->res_dual > T->tol_feas * 100.0 && I->res_dual > S->prev_res_dual * 100.0) ||
                (I->res_primal > T->tol_feas * 100.0 && I->res_primal > S->prev_res_primal * 100.0))
                I->status = ST_INSUFFICIENT_PROGRESS;
        }
    }
    if (I->status == ST_UNSOLVED) {
        if (T->max_iter == I->iterations) I->status = ST_MAX_ITERATIONS;
        else if (I->solve_time > T->time_limit) I->status = ST_MAX_TIME;
    }
    return I->status != ST_UNSOLVED;
}

static int kktsystem_solve_constant_rhs(oipm_t *S)
{
    for (idx i = 0; i < S->n; i++) S->workx[i] = -1.0 * S->q[i] + 0.0 * S->workx[i];
    kkt_setrhs(S, S->workx, S->b);
    return kkt_solve(S, S->x2, S->z2);
}
static int kktsystem_update(oipm_t *S)
{
    double t = now_s();
    int ok = kkt_update(S);
    if (ok) ok = kktsystem_solve_constant_rhs(S);
    S->info.t_kkt_update += now_s() - t;
    return ok;
}

/* lhs <- solution of the reduced Newton system for rhs (step_rhs) */
static int kktsystem_solve(oipm_t *S, int combined)
{
    double t0 = now_s();
    idx n = S->n, m = S->m;
    double *workx = S->workx, *workz = S->workz, *dsc = S->work_conic;
    memcpy(workx, S->rx_, (size_t)n * sizeof(double));
    if (!combined) memcpy(dsc, S->vs, (size_t)m * sizeof(double));
    else cones_ds_from_dz_offset(S, dsc, S->rs_, S->vz);
    for (idx i = 0; i < m; i++) workz[i] = 1.0 * dsc[i] + -1.0 * S->rz_[i];
    kkt_setrhs(S, workx, workz);
    int ok = kkt_solve(S, S->x1, S->z1);
    if (!ok) { S->info.t_kkt_solve += now_s() - t0; return 0; }
    double *xi = workx;
    double tinv = 1.0 / S->vtau;
    for (idx i = 0; i < n; i++) xi[i] = tinv * S->vx[i] + 0.0 * xi[i];
    double tau_num = S->rtau_ - S->rkap_ / S->vtau + vdot(S->q, S->x1, n) + vdot(S->b, S->z1, m) +
                     2.0 * quad_form_triu(&S->P, xi, S->x1);
    for (idx i = 0; i < n; i++) xi[i] = -1.0 * S->x2[i] + 1.0 * xi[i];
    double tau_den = S->vkap / S->vtau - vdot(S->q, S->x2, n) - vdot(S->b, S->z2, m);
    tau_den += quad_form_triu(&S->P, xi, xi) - quad_form_triu(&S->P, S->x2, S->x2);
    S->ltau = tau_num / tau_den;
    for (idx i = 0; i < n; i++) S->lx[i] = 1.0 * S->x1[i] + S->ltau * S->x2[i];
    for (idx i = 0; i < m; i++) S->lz[i] = 1.0 * S->z1[i] + S->ltau * S->z2[i];
    cones_mul_Hs(S, S->ls, S->lz);
    for (idx i = 0; i < m; i++) S->ls[i] = -1.0 * dsc[i] + -1.0 * S->ls[i];
    S->lkap = -(S->rkap_ + S->vkap * S->ltau) / S->vtau;
    S->info.t_kkt_solve += now_s() - t0;
    return 1;
}

static int solve_initial_point(oipm_t *S)
{
    idx n = S->n, m = S->m; int ok;
    if (S->P.colptr[n] == 0) {
        for (idx i = 0; i < n; i++) S->workx[i] = 0.0;
        memcpy(S->workz, S->b, (size_t)m * sizeof(double));
        kkt_setrhs(S, S->workx, S->workz);
        ok = kkt_solve(S, S->vx, S->vs);
        for (idx i = 0; i < m; i++) S->vs[i] = -S->vs[i];
        if (!ok) return ok;
        for (idx i = 0; i < n; i++) S->workx[i] = -1.0 * S->q[i] + 0.0 * S->workx[i];
        for (idx i = 0; i < m; i++) S->workz[i] = 0.0;
        kkt_setrhs(S, S->workx, S->workz);
        ok = kkt_solve(S, NULL, S->vz);
    } else {
        for (idx i = 0; i < n; i++) S->workx[i] = -S->q[i];
        memcpy(S->workz, S->b, (size_t)m * sizeof(double));
        kkt_setrhs(S, S->workx, S->workz);
        ok = kkt_solve(S, S->vx, S->vz);
        for (idx i = 0; i < m; i++) S->vs[i] = -S->vz[i];
    }
    return ok;
}

static double calc_step_length(oipm_t *S, int combined)
{
    double at = S->ltau < 0.0 ? -S->vtau / S->ltau : 1.7976931348623157e308;
    double ak = S->lkap < 0.0 ? -S->vkap / S->lkap : 1.7976931348623157e308;
    double a = fmin(fmin(at, ak), 1.0);
    a = cones_step_length(S, S->lz, S->ls, S->vz, S->vs, a);
    if (combined) a *= S->set.max_step_fraction;
    return a;
}

/* variables.rs:205-228 */
static double variables_barrier(oipm_t *S, double a)
{
    double central_coef = (double)(S->degree + 1);
    double cur_tau = S->vtau + a * S->ltau, cur_kap = S->vkap + a * S->lkap;
    double sz = 0.0;
    for (idx i = 0; i < S->m; i++) { double si = S->vs[i] + a * S->ls[i], zi = S->vz[i] + a * S->lz[i]; sz += si * zi; }
    double mu = (sz + cur_tau * cur_kap) / central_coef;
    double barrier = central_coef * logsafe(mu) - logsafe(cur_tau) - logsafe(cur_kap);
    barrier += cones_compute_barrier(S, S->vz, S->vs, S->lz, S->ls, a);
    return barrier;
}
/* core/solver.rs:548-584 */
static double get_step_length(oipm_t *S, int combined, int scaling)
{
    double a = calc_step_length(S, combined);
    if (!S->all_symmetric && combined && scaling == SCALING_DUAL) {
        double step = S->set.linesearch_backtrack_step;
        for (int it = 0; it < 50; it++) {
            double barrier = variables_barrier(S, a);
            if (barrier < 1.0) return a;
            a = step * a;
        }
    }
    return a;
}

static void save_prev(oipm_t *S)
{
    oipm_info *I = &S->info;
    S->prev_cost_primal = I->cost_primal; S->prev_cost_dual = I->cost_dual;
    S->prev_res_primal = I->res_primal; S->prev_res_dual = I->res_dual;
    S->prev_gap_abs = I->gap_abs; S->prev_gap_rel = I->gap_rel;
    memcpy(S->px, S->vx, (size_t)S->n * sizeof(double)); memcpy(S->ps, S->vs, (size_t)S->m * sizeof(double));
    memcpy(S->pz, S->vz, (size_t)S->m * sizeof(double)); S->ptau = S->vtau; S->pkap = S->vkap;
}
static void reset_to_prev(oipm_t *S)
{
    oipm_info *I = &S->info;
    I->cost_primal = S->prev_cost_primal; I->cost_dual = S->prev_cost_dual;
    I->res_primal = S->prev_res_primal; I->res_dual = S->prev_res_dual;
    I->gap_abs = S->prev_gap_abs; I->gap_rel = S->prev_gap_rel;
    memcpy(S->vx, S->px, (size_t)S->n * sizeof(double)); memcpy(S->vs, S->ps, (size_t)S->m * sizeof(double));
    memcpy(S->vz, S->pz, (size_t)S->m * sizeof(double)); S->vtau = S->ptau; S->vkap = S->pkap;
}

/* per-iteration trace for the parity tests: [mu, alpha, sigma, res_primal, res_dual, gap_abs] */
#define OIPM_TRACE_W 6
int oipm_solve(oipm_t *S, double *trace, int32_t trace_cap)
{
    if (!S->ldl) return -1;
    idx n = S->n, m = S->m;
    oipm_info *I = &S->info;
    int iter = 0; double sigma = 1.0, alpha = 0.0, mu = 0.0;
    double t0 = now_s();
    I->status = ST_UNSOLVED; I->iterations = 0;
    I->t_kkt_update = I->t_kkt_solve = I->t_scale_cones = 0; I->n_refactor = I->n_ldl_solve = 0;
    /* default start (core/solver.rs:525-541) */
    if (S->all_symmetric) {
        cones_set_identity(S);
        kktsystem_update(S);
        solve_initial_point(S);
        shift_to_cone_interior(S, S->vs, 1);
        shift_to_cone_interior(S, S->vz, 0);
    } else {
        cones_unit_initialization(S, S->vz, S->vs);     /* variables.rs:173-179 */
        for (idx i = 0; i < n; i++) S->vx[i] = 0.0;
    }
    S->vtau = 1.0; S->vkap = 1.0;
    /* core/solver.rs:277-280: the dual-only scaling from the start when a cone (GenPow) has no primal-dual one */
    int scaling = S->allows_primal_dual ? SCALING_PRIMAL_DUAL : SCALING_DUAL;

    for (;;) {
        residuals_update(S);
        mu = (S->dot_sz + S->vtau * S->vkap) / (double)(S->degree + 1);
        I->mu = mu; I->step_length = alpha; I->sigma = sigma; I->iterations = iter;
        info_update(S, t0);
        if (trace && iter < trace_cap) {
            double *tr = trace + (size_t)iter * OIPM_TRACE_W;
            tr[0] = mu; tr[1] = alpha; tr[2] = sigma; tr[3] = I->res_primal; tr[4] = I->res_dual; tr[5] = I->gap_abs;
        }
        if (check_termination(S, iter)) {
            if (getenv("OIPM_DEBUG")) fprintf(stderr, "[oipm] terminated status %d at iter %d (gap_abs %g gap_rel %g pres %g dres %g kt %g)\n", I->status, iter, I->gap_abs, I->gap_rel, I->res_primal, I->res_dual, I->ktratio);
            /* strategy_checkpoint_insufficient_progress (core/solver.rs:586-611) */
            if (I->status != ST_INSUFFICIENT_PROGRESS) break;
            reset_to_prev(S);
            if (!S->all_symmetric && scaling == SCALING_PRIMAL_DUAL) { I->status = ST_UNSOLVED; scaling = SCALING_DUAL; continue; }
            break;
        }
        double ts = now_s();
        int okscale = cones_update_scaling(S, S->vs, S->vz, mu, scaling);
        I->t_scale_cones += now_s() - ts;
        if (!okscale) { if (getenv("OIPM_DEBUG")) fprintf(stderr, "[oipm] scaling failed at iter %d\n", iter); I->status = ST_NUMERICAL_ERROR; break; }
        iter += 1;
        int ok = kktsystem_update(S);
        /* affine rhs */
        memcpy(S->rx_, S->rx, (size_t)n * sizeof(double));
        memcpy(S->rz_, S->rz, (size_t)m * sizeof(double));
        cones_affine_ds(S, S->rs_, S->vs);
        S->rtau_ = S->rtau; S->rkap_ = S->vtau * S->vkap;
        ok = ok && kktsystem_solve(S, 0);
        if (ok) {
            alpha = get_step_length(S, 0, scaling);
            sigma = (1.0 - alpha) * (1.0 - alpha) * (1.0 - alpha);
            double mm = iter > 1 ? 1.0 : alpha;
            double dsm = sigma * mu;
            for (idx i = 0; i < n; i++) S->rx_[i] = (1.0 - sigma) * S->rx[i] + 0.0 * S->rx_[i];
            S->rtau_ = (1.0 - sigma) * S->rtau;
            S->rkap_ = -dsm + mm * S->ltau * S->lkap + S->vtau * S->vkap;
            if (mm != 1.0) for (idx i = 0; i < m; i++) S->lz[i] *= mm;
            cones_combined_ds_shift(S, S->rz_, S->lz, S->ls, dsm);
            for (idx i = 0; i < m; i++) S->rs_[i] = 1.0 * S->rz_[i] + 1.0 * S->rs_[i];
            for (idx i = 0; i < m; i++) S->rz_[i] = (1.0 - sigma) * S->rz[i] + 0.0 * S->rz_[i];
            ok = kktsystem_solve(S, 1);
        }
        /* strategy_checkpoint_numerical_error (core/solver.rs:613-631) */
        if (!ok && !S->all_symmetric && scaling == SCALING_PRIMAL_DUAL) { alpha = 0.0; scaling = SCALING_DUAL; continue; }
        if (!ok) { if (getenv("OIPM_DEBUG")) fprintf(stderr, "[oipm] kkt failure at iter %d\n", iter); I->status = ST_NUMERICAL_ERROR; alpha = 0.0; break; }
        alpha = get_step_length(S, 1, scaling);
        /* strategy_checkpoint_small_step (core/solver.rs:633-651) */
        if (!S->all_symmetric && scaling == SCALING_PRIMAL_DUAL && alpha < S->set.min_switch_step_length) { alpha = 0.0; scaling = SCALING_DUAL; continue; }
        if (alpha <= fmax(0.0, S->set.min_terminate_step_length)) { if (getenv("OIPM_DEBUG")) fprintf(stderr, "[oipm] small step %g at iter %d\n", alpha, iter); I->status = ST_INSUFFICIENT_PROGRESS; alpha = 0.0; break; }
        save_prev(S);
        for (idx i = 0; i < n; i++) S->vx[i] = alpha * S->lx[i] + 1.0 * S->vx[i];
        for (idx i = 0; i < m; i++) S->vs[i] = alpha * S->ls[i] + 1.0 * S->vs[i];
        for (idx i = 0; i < m; i++) S->vz[i] = alpha * S->lz[i] + 1.0 * S->vz[i];
        S->vtau += alpha * S->ltau; S->vkap += alpha * S->lkap;
    }
    if (alpha == 0.0) { I->mu = mu; I->step_length = alpha; I->sigma = sigma; I->iterations = iter; }
    /* post-process: "almost" statuses after an error / limit exit (info.rs:95-105) */
    if (I->status == ST_NUMERICAL_ERROR || I->status == ST_INSUFFICIENT_PROGRESS ||
        I->status == ST_MAX_ITERATIONS || I->status == ST_MAX_TIME) {
        const oipm_settings *T = &S->set;
        check_convergence(S, T->reduced_tol_gap_abs, T->reduced_tol_gap_rel, T->reduced_tol_feas,
                          T->reduced_tol_infeas_abs, T->reduced_tol_infeas_rel, T->reduced_tol_ktratio,
                          ST_ALMOST_SOLVED, ST_ALMOST_PRIMAL_INFEASIBLE, ST_ALMOST_DUAL_INFEASIBLE);
    }
    I->solve_time = now_s() - t0;
    return 0;
}

/* unscaled solution (variables.rs:262-285, solution.rs:68-111) */
void oipm_get_solution(oipm_t *S, double *x, double *z, double *s, double *obj, double *obj_dual)
{
    int st = S->info.status;
    int infeas = (st == ST_PRIMAL_INFEASIBLE || st == ST_DUAL_INFEASIBLE ||
                  st == ST_ALMOST_PRIMAL_INFEASIBLE || st == ST_ALMOST_DUAL_INFEASIBLE);
    double scaleinv = infeas ? 1.0 / S->vkap : 1.0 / S->vtau;
    double cinv = 1.0 / S->c;
    for (idx i = 0; i < S->n; i++) x[i] = S->vx[i] * S->d[i] * scaleinv;
    if (!S->keep) {
        for (idx i = 0; i < S->m; i++) z[i] = S->vz[i] * S->e[i] * (scaleinv * cinv);
        for (idx i = 0; i < S->m; i++) s[i] = S->vs[i] * S->einv[i] * scaleinv;
    } else {   /* reverse_presolve (presolver.rs:127-150): dropped rows get s = infinity bound, z = 0 */
        idx c = 0;
        for (idx i = 0; i < S->mfull; i++) {
            if (S->keep[i]) { z[i] = S->vz[c] * S->e[c] * (scaleinv * cinv); s[i] = S->vs[c] * S->einv[c] * scaleinv; c++; }
            else { z[i] = 0.0; s[i] = g_infinity; }
        }
    }
    *obj = infeas ? NAN : S->info.cost_primal;
    *obj_dual = infeas ? NAN : S->info.cost_dual;
}
void oipm_get_info(const oipm_t *S, oipm_info *out) { *out = S->info; }

/* expose single pieces for unit-level parity tests of the CUDA cone kernels */
int oipm_test_update_scaling(oipm_t *S, const double *s, const double *z) { return cones_update_scaling(S, s, z, 0.0, SCALING_PRIMAL_DUAL); }
int oipm_test_update_scaling_ex(oipm_t *S, const double *s, const double *z, double mu, int strategy) { return cones_update_scaling(S, s, z, mu, strategy); }
double oipm_test_compute_barrier(oipm_t *S, const double *z, const double *s, const double *dz, const double *ds, double a) { return cones_compute_barrier(S, z, s, dz, ds, a); }
void oipm_test_unit_initialization(oipm_t *S, double *z, double *s) { cones_unit_initialization(S, z, s); }
double oipm_test_wright_omega(double z) { return wright_omega(z); }
/* per-cone state after update_scaling: out = [H_dual(6), Hs(6), grad(3), z(3)] */
void oipm_test_ns3_state(const oipm_t *S, idx k, double *out)
{
    const ns3_t *K = S->cones[k].ns;
    if (!K) return;
    for (int i = 0; i < 6; i++) { out[i] = K->H_dual[i]; out[6 + i] = K->Hs[i]; }
    for (int i = 0; i < 3; i++) { out[12 + i] = K->grad[i]; out[15 + i] = K->z[i]; }
}
void oipm_test_get_Hs(oipm_t *S, double *Hs) { cones_get_Hs(S, Hs); }
idx oipm_nHs(const oipm_t *S) { return S->nHs; }
void oipm_test_mul_Hs(oipm_t *S, double *y, const double *x) { cones_mul_Hs(S, y, x); }
void oipm_test_affine_ds(oipm_t *S, double *ds) { cones_affine_ds(S, ds, S->vs); }
void oipm_test_affine_ds_ex(oipm_t *S, double *ds, const double *s) { cones_affine_ds(S, ds, s); }
void oipm_test_combined_ds_shift(oipm_t *S, double *shift, double *sz, double *ss, double sm) { cones_combined_ds_shift(S, shift, sz, ss, sm); }
void oipm_test_ds_from_dz_offset(oipm_t *S, double *out, const double *ds, const double *z) { cones_ds_from_dz_offset(S, out, ds, z); }
double oipm_test_step_length(oipm_t *S, const double *dz, const double *ds, const double *z, const double *s, double amax) { return cones_step_length(S, dz, ds, z, s, amax); }


/* ---- test hooks for the algebra kernels of SURVEY 8 rows a11 / a20, pinned on the reference's own known answers
 * (src/algebra/tests/matrix.rs: test_gemv, test_symv, test_quad_form; src/algebra/tests/vector.rs: test_norm*,
 * test_dot) by tests/test_oracle_algebra.py ---- */
static csc hook_csc(idx m, idx n, const idx *cp, const idx *rv, const double *nz)
{ csc A; A.m = m; A.n = n; A.colptr = (idx *)cp; A.rowval = (idx *)rv; A.nzval = (double *)nz; return A; }
void oipm_test_symv(idx n, const idx *cp, const idx *rv, const double *nz, double *y, const double *x, double a, double b)
{ csc A = hook_csc(n, n, cp, rv, nz); symv_tri(&A, y, x, a, b); }
double oipm_test_quad_form(idx n, const idx *cp, const idx *rv, const double *nz, const double *y, const double *x)
{ csc A = hook_csc(n, n, cp, rv, nz); return quad_form_triu(&A, y, x); }
void oipm_test_gemv(idx m, idx n, const idx *cp, const idx *rv, const double *nz, int transposed, double *y, const double *x,
                    double a, double b)
{ csc A = hook_csc(m, n, cp, rv, nz); if (transposed) gemv_T(&A, y, x, a, b); else gemv_N(&A, y, x, a, b); }
double oipm_test_vec(int what, const double *x, const double *v, idx n)
{ return what == 0 ? vnorm(x, n) : what == 1 ? vnorm_inf(x, n) : what == 2 ? vnorm_scaled(x, v, n) : vdot(x, v, n); }
