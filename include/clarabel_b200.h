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

/* timing helper for benches: runs `reps` refactors (or solves) back to back
 * on the device and returns the average milliseconds measured with CUDA
 * events on the handle's stream. */
double cldl_time_refactor_ms(cldl_t *h, int reps);
double cldl_time_solve_ms(cldl_t *h, int reps);

#ifdef __cplusplus
}
#endif
#endif
