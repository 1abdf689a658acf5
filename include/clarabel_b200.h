/*
 * clarabel_b200.h -- C-ABI of the B200-native KKT backend for Clarabel-style
 * interior point solvers.  Plain pointers and sizes only; no C++/torch types.
 *
 * LEVEL 1  (cldl_*)  replaces the reference's `DirectLDLSolver` plugin trait
 *   /root/reference/src/solver/core/kktsolvers/direct/quasidef/mod.rs:14-26
 *   and its qdldl adapter .../ldlsolvers/qdldl.rs:19-107.  One handle = one
 *   factorisation object living on one GPU.
 *
 * Conventions (follow the reference's existing C-ABI, src/julia/interface.rs):
 *   - opaque handle + explicit destroy;
 *   - index type is uint64_t (Rust `usize`);
 *   - the two trait methods that return `bool` return 1 (true) / 0 (false);
 *     everything else returns 0 on success and a negative CLDL_E_* code on
 *     failure.  Nothing unwinds or aborts.
 *   - every entry point selects the handle's device itself (cudaSetDevice),
 *     so a handle may be used from a thread other than its creator
 *     (directldlkktsolver.rs:13-16: the trait object is Send + Sync);
 *   - `_dev` twins take DEVICE pointers (same meaning otherwise) and enqueue
 *     on the handle's stream without synchronising.
 */
#ifndef CLARABEL_B200_H
#define CLARABEL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cldl_handle cldl_t;

enum {
  CLDL_OK = 0,
  CLDL_E_DIM = -1,          /* QDLDLError::IncompatibleDimension */
  CLDL_E_EMPTY_COLUMN = -2, /* QDLDLError::EmptyColumn          */
  CLDL_E_NOT_TRIU = -3,     /* QDLDLError::NotUpperTriangular   */
  CLDL_E_ZERO_PIVOT = -4,   /* QDLDLError::ZeroPivot            */
  CLDL_E_BAD_PERM = -5,     /* QDLDLError::InvalidPermutation   */
  CLDL_E_CUDA = -20,        /* no device / CUDA runtime failure */
  CLDL_E_ARG = -21,
  CLDL_E_NOT_FACTORED = -22 /* solve() before refactor(): the reference panics (qdldl.rs:118) */
};

enum { CLDL_ORDER_AMD = 1, CLDL_ORDER_ND = 2, CLDL_ORDER_BEST = 3 };

/* Options read at construction.  Mirrors the fields of CoreSettings that the
 * qdldl adapter forwards (ldlsolvers/qdldl.rs:35-42) plus device selection. */
typedef struct {
  double regularize_eps;    /* settings.dynamic_regularization_eps   (default 1e-13) */
  double regularize_delta;  /* settings.dynamic_regularization_delta (default 2e-7)  */
  int32_t regularize_enable;/* adapter always passes true (qdldl.rs:38)               */
  double amd_dense_scale;   /* 1.5 in the reference adapter (qdldl.rs:41)             */
  int32_t ordering;         /* CLDL_ORDER_*; ignored when a permutation is supplied   */
  int32_t device;           /* CUDA device ordinal                                    */
  int32_t max_panel;        /* 0 = default                                            */
  int32_t nd_leaf;          /* 0 = default                                            */
  /* one factorisation on several GPUs: this handle is rank `shard_rank` of `shard_nranks`; it factors
   * and solves the subtrees it owns plus the replicated top of the assembly tree.  0 / 1 ranks = whole tree. */
  int32_t shard_nranks, shard_rank;
} cldl_opts;

/* LinearSolverInfo (kktsolvers/mod.rs:24-38) + factorisation counters
 * (qdldl.rs:104-112). */
typedef struct {
  char name[16];            /* "cudaldl" */
  uint32_t threads;         /* resident device threads used per launch wave; 0 = n/a */
  int32_t direct;           /* 1 */
  uint64_t nnzA;
  uint64_t nnzL;            /* entries of L the reference would report (simplicial count) */
  uint64_t nnzL_stored;     /* entries actually stored in dense supernodal panels */
  uint64_t regularize_count;
  uint64_t positive_inertia;
  uint64_t n_supernodes;
  uint64_t n_levels;
  double flops;             /* dense flops per numeric factorisation */
  int32_t ordering_used;    /* 0 = caller's permutation */
} cldl_info_t;

void cldl_default_opts(cldl_opts *o);

/* Constructor: DirectLDLSolver ctor signature of ldlsolvers/config.rs:19-20,
 * `fn(&CscMatrix<T>, &[i8] Dsigns, &CoreSettings<T>, Option<Vec<usize>> perm)`.
 * (colptr,rowval,nzval) is the n x n upper-triangular KKT matrix in CSC with a
 * structural entry on every diagonal.  Performs ordering + symbolic analysis on
 * the host and uploads the static maps; like the reference adapter it does NOT
 * produce numeric factors ("logical" factorisation) -- call cldl_refactor. */
int cldl_create(cldl_t **out, uint64_t n, const uint64_t *colptr, const uint64_t *rowval,
                const double *nzval, const int8_t *dsigns, const cldl_opts *opts,
                const uint64_t *perm_or_null);
void cldl_destroy(cldl_t *h);

/* DirectLDLSolver::update_values / scale_values / offset_values
 * (mod.rs:15-17, qdldl.rs:142-183).  `index` addresses entries of the
 * caller's nzval array. */
int cldl_update_values(cldl_t *h, const uint64_t *index, const double *values, uint64_t len);
int cldl_scale_values(cldl_t *h, const uint64_t *index, uint64_t len, double scale);
int cldl_offset_values(cldl_t *h, const uint64_t *index, uint64_t len, double offset,
                       const int8_t *signs);

/* DirectLDLSolver::refactor (mod.rs:19) -> bool: all reciprocal pivots finite
 * (ldlsolvers/qdldl.rs:99-106).  Returns 1/0, or CLDL_E_ZERO_PIVOT when an
 * exact zero pivot is met with regularisation disabled (qdldl.rs:527,656). */
int cldl_refactor(cldl_t *h);

/* DirectLDLSolver::solve (mod.rs:18): x <- K^{-1} b; b is left untouched
 * (ldlsolvers/qdldl.rs:93-97).  Host buffers of length n. */
int cldl_solve(cldl_t *h, double *x, const double *b);

void cldl_info(const cldl_t *h, cldl_info_t *info);

/* The permutation actually used (new k <- old perm[k]); the parity tests hand
 * it to the CPU oracle so both sides eliminate in the same order. */
int cldl_get_perm(const cldl_t *h, uint64_t *perm_out);

/* ---- device-pointer twins (asynchronous on the handle's stream) ---- */
int cldl_update_values_dev(cldl_t *h, const int32_t *d_index, const double *d_values, uint64_t len);
int cldl_set_values_dev(cldl_t *h, const double *d_nzval);   /* whole array, caller order */
int cldl_refactor_dev(cldl_t *h);                            /* enqueue only; status via cldl_sync_status */
int cldl_solve_dev(cldl_t *h, double *d_x, const double *d_b);
int cldl_sync_status(cldl_t *h);                             /* sync + refactor verdict (1/0/neg) */
void *cldl_stream(cldl_t *h);                                /* cudaStream_t */
double *cldl_values_dev(cldl_t *h);                          /* device copy of nzval, caller order */

/* ---- one factorisation on several GPUs (cldl_opts.shard_nranks > 1) ----
 * Every rank creates its handle from the same matrix with its own shard_rank; the symbolic analysis and the
 * subtree-to-rank plan are deterministic, so all ranks agree on them.  A rank factors / solves the subtrees it owns
 * and the replicated top of the assembly tree; between the two phases the caller moves the packed contributions of
 * every rank to every other rank (NCCL all-gather / broadcast over NVLink, or plain copies when the handles share a
 * device) -- the library only packs and unpacks DEVICE buffers:
 *   what = 0  update matrices of the rank's cut roots   (between the refactor phases)
 *   what = 1  update vectors of the rank's cut roots    (between the solve phases)
 *   what = 2  the x entries the rank computed           (after solve phase 1: the all-gather of the solution)
 * cldl_shard_count(h, what, r) = doubles rank r contributes.  cldl_refactor / cldl_solve refuse on such a handle.
 *   refactor:  phase 0 on every rank -> pack(0) -> exchange -> unpack(0, r) for r != me -> phase 1 -> cldl_sync_status
 *   solve:     phase 0 -> pack(1) -> exchange -> unpack(1, r) -> phase 1 -> pack(2, x) -> exchange -> unpack(2, r, x)
 * cldl_shard_counts: {regularize_count, positive_inertia} of the owned phase, then of owned + top (after
 * cldl_sync_status); the global count is sum_r owned_r + (total_0 - owned_0). */
int cldl_shard_refactor_phase_dev(cldl_t *h, int phase);
int cldl_shard_solve_phase_dev(cldl_t *h, double *d_x, const double *d_b, int phase);
uint64_t cldl_shard_count(const cldl_t *h, int what, int rank);
int cldl_shard_pack_dev(cldl_t *h, int what, double *d_buf, const double *d_x);
int cldl_shard_unpack_dev(cldl_t *h, int what, int rank, const double *d_buf, double *d_x);
int cldl_shard_counts(const cldl_t *h, uint64_t *out4);
/* Transport for a sharded handle: an all-gather of `count` doubles per rank between DEVICE buffers (d_send: count
 * doubles of this rank; d_recv: nranks * count doubles, rank r's block at r * count).  It is called from the host
 * between the phases, after the send buffer is complete, and must return once d_recv is usable (0 = ok).  With a
 * transport installed, cldl_refactor(_dev) / cldl_solve(_dev) and the whole cipm_* driver run their phases and
 * exchanges themselves (contributions are padded to the largest one); every rank then executes the same interior
 * point iterations on identical data and only the factorisation / triangular solves are split. */
typedef int (*cldl_allgather_fn)(void *ctx, const double *d_send, double *d_recv, uint64_t count);
int cldl_set_transport(cldl_t *h, cldl_allgather_fn fn, void *ctx);
/* The same exchanges as stream-ordered NCCL all-gathers issued by the library itself on the handle's stream (no host
 * synchronisation between pack, collective and unpack).  The library is not linked against NCCL: `libpath` names the
 * libnccl.so.2 the calling process has loaded already (two NCCL builds in one process do not mix; NULL = the loader's
 * default).  Rank 0 draws the 128-byte unique id and the binding broadcasts it; cldl_set_nccl is collective (it
 * calls ncclCommInitRank).  Takes precedence over a callback transport. */
int cldl_nccl_unique_id(const char *libpath, unsigned char *id128);
int cldl_set_nccl(cldl_t *h, const char *libpath, const unsigned char *id128, int nranks, int rank);
int cldl_copy_dev(void *d_dst, const void *d_src, uint64_t bytes);   /* device-to-device copy, for transports in bindings */

/* timing helper for benches: runs `reps` refactors (or solves) back to back
 * on the device and returns the average milliseconds measured with CUDA
 * events on the handle's stream. */
double cldl_time_refactor_ms(cldl_t *h, int reps);
double cldl_time_solve_ms(cldl_t *h, int reps);


/* ======================================================================
 * LEVEL 2  (cipm_* / ckkt_* / ccone_*)  device-resident KKT system, cone
 * engine and interior-point driver.  Replaces, for the symmetric cones
 * (Zero / Nonnegative / SecondOrder):
 *   trait KKTSolver      src/solver/core/kktsolvers/mod.rs:7-19
 *     + DirectLDLKKTSolver .../direct/quasidef/directldlkktsolver.rs:18-405
 *   trait Cone / CompositeCone  src/solver/core/cones/mod.rs:42-154,
 *                               compositecone.rs:197-352
 *   DefaultSolver::new / IPSolver::solve
 *                        src/solver/implementations/default/solver.rs:57-126,
 *                        src/solver/core/solver.rs:224-465
 * One handle owns the equilibrated problem, the cone set, the KKT matrix and
 * its LDL^T on one GPU.  Vectors never leave the device during a solve.
 * ==================================================================== */
typedef struct cipm_handle cipm_t;

/* SupportedConeT tags (supportedcone.rs:17-52).  ExponentialConeT() and PowerConeT(alpha) occupy 3 rows each;
 * the exponent of a power cone travels in cone_params (cipm_create_ex); GenPowerConeT(alpha, dim2) has
 * cone_dims = len(alpha) and its dim2 / exponents in the two extra arrays of cipm_create_gp.
 * PSDTriangleConeT(n): cone_dims = n (matrix dimension, n (n + 1) / 2 rows); n <= 128, larger cones are refused. */
enum { CIPM_CONE_ZERO = 0, CIPM_CONE_NONNEG = 1, CIPM_CONE_SOC = 2, CIPM_CONE_PSD = 3, CIPM_CONE_EXP = 4,
       CIPM_CONE_POW = 5, CIPM_CONE_GENPOW = 6 };
/* ScalingStrategy (src/solver/core/cones/mod.rs) */
enum { CIPM_SCALING_PRIMAL_DUAL = 0, CIPM_SCALING_DUAL = 1 };

/* SolverStatus (src/solver/core/traits.rs / default/info.rs) */
enum { CIPM_UNSOLVED = 0, CIPM_SOLVED, CIPM_PRIMAL_INFEASIBLE, CIPM_DUAL_INFEASIBLE, CIPM_ALMOST_SOLVED,
       CIPM_ALMOST_PRIMAL_INFEASIBLE, CIPM_ALMOST_DUAL_INFEASIBLE, CIPM_MAX_ITERATIONS, CIPM_MAX_TIME,
       CIPM_NUMERICAL_ERROR, CIPM_INSUFFICIENT_PROGRESS };

/* The fields of DefaultSettings the path reads (default/settings.rs:30-193),
 * same names, same defaults (cipm_default_settings). */
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
  int32_t presolve_enable;   /* drop nonnegative rows with an infinite bound (presolver.rs); default 1 */
} cipm_settings;

/* DefaultInfo (default/info.rs:13-64) + timers of core/solver.rs:330-396 + counters */
typedef struct {
  int32_t status;
  uint32_t iterations;
  double cost_primal, cost_dual, res_primal, res_dual, res_primal_inf, res_dual_inf;
  double gap_abs, gap_rel, ktratio, mu, step_length, sigma;
  double solve_time;          /* host wall clock, seconds */
  double device_ms;           /* CUDA events around the whole solve on the handle's stream */
  double t_kkt_update, t_kkt_solve, t_scale_cones;   /* "kkt update" / "kkt solve" / "scale cones" */
  uint64_t n_refactor, n_ldl_solve, n_ir_steps, regularize_count;
  uint64_t nnzK, nnzL, kkt_dim;
} cipm_info;

void cipm_default_settings(cipm_settings *s);

/* DefaultSolver::new(P, q, A, b, cones, settings).  P: n x n upper triangle CSC;
 * A: m x n CSC; cones: parallel arrays (type tag, dimension).  Cones are
 * collapsed, data equilibrated (Ruiz), KKT assembled and analysed here.
 * kkt_perm_or_null: optional elimination order for the (n+m+p) KKT system. */
int cipm_create(cipm_t **out, uint64_t n, uint64_t m, const uint64_t *P_colptr, const uint64_t *P_rowval,
                const double *P_nzval, const double *q, const uint64_t *A_colptr, const uint64_t *A_rowval,
                const double *A_nzval, const double *b, uint64_t ncones, const int32_t *cone_types,
                const uint64_t *cone_dims, const cipm_settings *settings, const cldl_opts *ldl_opts,
                const uint64_t *kkt_perm_or_null);
/* Same, with one double per cone: the exponent alpha of a CIPM_CONE_POW entry (PowerConeT(alpha),
 * supportedcone.rs:36-38; the Julia interface carries it the same way, julia/types.rs:19-26), ignored for the other
 * cone types.  cone_params may be NULL when the problem has no power cones. */
int cipm_create_ex(cipm_t **out, uint64_t n, uint64_t m, const uint64_t *P_colptr, const uint64_t *P_rowval,
                   const double *P_nzval, const double *q, const uint64_t *A_colptr, const uint64_t *A_rowval,
                   const double *A_nzval, const double *b, uint64_t ncones, const int32_t *cone_types,
                   const uint64_t *cone_dims, const double *cone_params, const cipm_settings *settings,
                   const cldl_opts *ldl_opts, const uint64_t *kkt_perm_or_null);
/* Same, with generalised power cones (GenPowerConeT(alpha, dim2), supportedcone.rs:44): for a CIPM_CONE_GENPOW entry
 * cone_dims[k] = len(alpha), genpow_dim2[k] = dim2 (ignored for other cones) and genpow_alpha holds the exponents of
 * all such cones concatenated in cone order (each set positive, summing to one). */
int cipm_create_gp(cipm_t **out, uint64_t n, uint64_t m, const uint64_t *P_colptr, const uint64_t *P_rowval,
                   const double *P_nzval, const double *q, const uint64_t *A_colptr, const uint64_t *A_rowval,
                   const double *A_nzval, const double *b, uint64_t ncones, const int32_t *cone_types,
                   const uint64_t *cone_dims, const double *cone_params, const uint64_t *genpow_dim2,
                   const double *genpow_alpha, const cipm_settings *settings, const cldl_opts *ldl_opts,
                   const uint64_t *kkt_perm_or_null);
/* installs the all-gather of a sharded factorisation (cldl_opts.shard_nranks > 1 in ldl_opts) on the solver's LDL */
int cipm_set_transport(cipm_t *h, cldl_allgather_fn fn, void *ctx);
int cipm_set_nccl(cipm_t *h, const char *libpath, const unsigned char *id128, int nranks, int rank);
uint64_t cipm_collective_count(const cipm_t *h);      /* NCCL all-gathers the handle has issued so far */
/* Solver::update_settings (core/solver.rs:207-211): new settings for the next cipm_solve; CLDL_E_ARG when a field that
 * only acts at construction differs (equilibration parameters, presolve_enable: settings.rs:307-335). */
int cipm_update_settings(cipm_t *h, const cipm_settings *settings);
void cipm_destroy(cipm_t *h);
int cipm_solve(cipm_t *h);                                   /* IPSolver::solve */
void cipm_get_info(const cipm_t *h, cipm_info *out);
int cipm_get_solution(cipm_t *h, double *x, double *z, double *s);   /* unscaled, host buffers */
uint64_t cipm_trace(const cipm_t *h, double *out, uint64_t cap_rows); /* rows of [mu,alpha,sigma,pres,dres,gap] */
/* device timestamps (ms since solve() start, CUDA events on the handle's stream) taken at the start of
 * every iteration; the last entry is the end of the solve. */
uint64_t cipm_iter_ms(const cipm_t *h, double *out, uint64_t cap);
uint64_t cipm_launch_count(void);
/* sizeof of {cldl_opts, cldl_info_t, cipm_settings, cipm_info} as compiled into the library: a binding checks its
 * own struct mirrors against them before the first call */
void cipm_abi_sizes(uint64_t *out4);
/* device-timed (CUDA events) average ms of: 0 numeric refactor, 1 one LDL solve, 2 one KKT solve incl. IR */
double cipm_time_ms(cipm_t *h, int which, int reps);                            /* kernels launched by this library so far */
/* Kernel-level test entry points: the sparse products and the reductions of the iteration body on caller data
 * (src/algebra/csc/matrix_math.rs:178-343, src/algebra/vecmath.rs:83-226), so that they can be compared with the
 * reference's own unit-test answers (src/algebra/tests/matrix.rs, vector.rs).
 * cipm_test_spmv: which = 0  y = a P x + b y (the handle's symmetric P), 1  y = a A x + b y, 2  y = a A' x + b y.
 * cipm_test_vec:  what = 0  ||x||_2, 1  ||x||_inf (NaN propagates), 2  ||x .* v||_2, 3  <x, v>. */
int cipm_test_spmv(cipm_t *h, int which, double *y, const double *x, double a, double b);
int cipm_test_vec(cipm_t *h, int what, const double *x, const double *v, uint64_t n, double *out);
/* get_infinity / set_infinity / default_infinity (src/src/utils/infbounds.rs; tests/presolve.rs:107-114): the
 * process-wide bound (default 1e20) beyond which a nonnegative-cone row counts as absent in the presolve; read when a
 * handle is created and when its solution is expanded. */
double cipm_get_infinity(void);
void cipm_set_infinity(double v);
void cipm_default_infinity(void);
uint64_t cipm_m_reduced(const cipm_t *h);   /* rows left after the inf-bound presolve (== m when nothing was dropped) */
/* DefaultProblemData::equilibration (problemdata.rs:229-312; pinned by tests/equilibration_bounds.rs): the Ruiz
 * scalings d [n], e [cipm_m_reduced] and the cost scaling c of the handle; any of the three pointers may be NULL. */
int cipm_get_equilibration(const cipm_t* h, double* d, double* e, double* c);
uint64_t cipm_kkt_dim(const cipm_t *h);
uint64_t cipm_kkt_nnz(const cipm_t *h);
int cipm_get_kkt(const cipm_t *h, uint64_t *colptr, uint64_t *rowval, double *nzval, int8_t *dsigns);
int cipm_get_kkt_perm(const cipm_t *h, uint64_t *perm);
void cipm_ldl_info(const cipm_t *h, cldl_info_t *info);

/* DefaultSolver::update_data (src/solver/implementations/default/data_updating.rs:68-163): overwrite the values of
 * P (triu, same pattern), q, A (same pattern), b in an existing solver; the stored equilibration is applied, symbolic
 * analysis and device plans are reused, the next cipm_solve starts from the default initial point.  NULL = unchanged. */
int cipm_update_data(cipm_t *h, const double *P_nzval, const double *q, const double *A_nzval, const double *b);

/* KKTSolver trait on the handle's KKT object (host buffers; x has length n, z length m).
 * ckkt_update / ckkt_solve return 1 (true) / 0 (false) like the trait's bools. */
int ckkt_update(cipm_t *h);                                  /* KKTSolver::update(cones, settings) */
int ckkt_setrhs(cipm_t *h, const double *rhsx, const double *rhsz);
int ckkt_solve(cipm_t *h, double *lhsx, double *lhsz);       /* LDL solve + iterative refinement */
int ckkt_update_P(cipm_t *h, const double *P_nzval_scaled);
int ckkt_update_A(cipm_t *h, const double *A_nzval_scaled);
int ckkt_get_values(cipm_t *h, double *nzval_out);           /* current (un-regularised) KKT values */

/* Cone trait on the handle's composite cone (host buffers of length m). */
int ccone_set_identity_scaling(cipm_t *h);
int ccone_update_scaling(cipm_t *h, const double *s, const double *z);   /* bool */
uint64_t ccone_Hs_len(const cipm_t *h);
int ccone_get_Hs(cipm_t *h, double *Hs);
int ccone_mul_Hs(cipm_t *h, double *y, const double *x);
int ccone_affine_ds(cipm_t *h, double *ds);
int ccone_combined_ds_shift(cipm_t *h, double *shift, const double *step_z, const double *step_s, double sigmamu);
int ccone_ds_from_dz_offset(cipm_t *h, double *out, const double *ds, const double *z);
int ccone_step_length(cipm_t *h, const double *dz, const double *ds, const double *z, const double *s,
                      double alpha_max, double *alpha_out);
int ccone_margins(cipm_t *h, const double *z, double *min_margin, double *pos_margin);
int ccone_scaled_unit_shift(cipm_t *h, double *z, double alpha, int primal);
/* the parts of the trait only nonsymmetric problems use (cones/mod.rs:61-66, 94-101, 148-153) */
int ccone_is_symmetric(const cipm_t *h);
int ccone_unit_initialization(cipm_t *h, double *z, double *s);
int ccone_update_scaling_ex(cipm_t *h, const double *s, const double *z, double mu, int strategy);   /* bool */
int ccone_affine_ds_ex(cipm_t *h, double *ds, const double *s);
int ccone_compute_barrier(cipm_t *h, const double *z, const double *s, const double *dz, const double *ds,
                          double alpha, double *barrier_out);

#ifdef __cplusplus
}
#endif
#endif
