// Host-only C entry points onto the symbolic analysis (no CUDA needed).
// Used by the CPU test-suite to validate orderings, supernodes, relative
// index maps and the level schedule, and by bench.py to report nnzL / flops.
#include <cstdint>
#include <cstring>
#include <vector>

#include "symbolic.h"

using cb::Symbolic;

extern "C" {

struct csym_handle { Symbolic S; };

// ordering: 0 = use perm (must be non-null), 1 AMD, 2 ND, 3 best-of
int csym_analyse(csym_handle** out, uint64_t n, const uint64_t* colptr, const uint64_t* rowval,
                 const uint64_t* perm_or_null, int ordering, double amd_dense_scale, int max_panel,
                 int nd_leaf) {
  *out = nullptr;
  if (n == 0 || n > 0x7fffffffu) return -1;
  std::vector<int64_t> Ap(n + 1);
  uint64_t nnz = colptr[n];
  std::vector<int32_t> Ai(nnz);
  for (uint64_t j = 0; j <= n; j++) Ap[j] = (int64_t)colptr[j];
  for (uint64_t p = 0; p < nnz; p++) { if (rowval[p] >= n) return -1; Ai[p] = (int32_t)rowval[p]; }
  std::vector<int> perm;
  if (perm_or_null) { perm.resize(n); for (uint64_t k = 0; k < n; k++) perm[k] = (int)perm_or_null[k]; }
  cb::SymbolicOptions so;
  so.ordering = ordering ? ordering : cb::ORDER_BEST;
  if (amd_dense_scale > 0) so.amd_dense_scale = amd_dense_scale;
  if (max_panel > 0) so.max_panel = max_panel;
  if (nd_leaf > 0) so.nd_leaf = nd_leaf;
  csym_handle* h = new csym_handle();
  int rc = cb::analyse((int)n, Ap.data(), Ai.data(), perm_or_null ? perm.data() : nullptr, so, h->S);
  if (rc) { delete h; return rc; }
  *out = h;
  return 0;
}
void csym_free(csym_handle* h) { delete h; }

// scalar getters: 0 n, 1 nsup, 2 nlevels, 3 nnzL_simplicial, 4 nnzL_stored, 5 upd_total,
// 6 ordering_used, 7 nnzA, 8 len(sn_rows), 9 len(child_list)
int64_t csym_scalar(const csym_handle* h, int which) {
  const Symbolic& S = h->S;
  switch (which) {
    case 0: return S.n;
    case 1: return S.nsup;
    case 2: return S.nlevels;
    case 3: return S.nnzL_simplicial;
    case 4: return S.nnzL_stored;
    case 5: return S.upd_total;
    case 6: return S.ordering_used;
    case 7: return S.nnzA;
    case 8: return (int64_t)S.sn_rows.size();
    case 9: return (int64_t)S.child_list.size();
    case 10: return S.L_alloc;
  }
  return -1;
}
double csym_flops(const csym_handle* h, int stored) { return stored ? h->S.flops_stored : h->S.flops_simplicial; }

// array getters copy into caller buffers (int64 for everything, for simplicity)
// which: 0 perm, 1 parent, 2 colcount, 3 sn_first, 4 sn_rowptr, 5 sn_rows, 6 sn_parent,
// 7 sn_level, 8 child_ptr, 9 child_list, 10 rel, 11 panel_off, 12 upd_off, 13 asm_ptr,
// 14 asm_src, 15 asm_dst, 16 level_ptr, 17 level_tasks, 18 iperm
int64_t csym_array(const csym_handle* h, int which, int64_t* out, int64_t cap) {
  const Symbolic& S = h->S;
  auto put = [&](auto const& v) -> int64_t {
    int64_t len = (int64_t)v.size();
    if (out) { if (cap < len) return -1; for (int64_t i = 0; i < len; i++) out[i] = (int64_t)v[i]; }
    return len;
  };
  switch (which) {
    case 0: return put(S.perm);
    case 1: return put(S.parent);
    case 2: return put(S.colcount);
    case 3: return put(S.sn_first);
    case 4: return put(S.sn_rowptr);
    case 5: return put(S.sn_rows);
    case 6: return put(S.sn_parent);
    case 7: return put(S.sn_level);
    case 8: return put(S.child_ptr);
    case 9: return put(S.child_list);
    case 10: return put(S.rel);
    case 11: return put(S.panel_off);
    case 12: return put(S.upd_off);
    case 13: return put(S.asm_ptr);
    case 14: return put(S.asm_src);
    case 15: return put(S.asm_dst);
    case 16: return put(S.level_ptr);
    case 17: return put(S.level_tasks);
    case 18: return put(S.iperm);
  }
  return -2;
}

// ordering for matrices with dense diagonal blocks (group[v] = block id or -1); returns the kind chosen
int csym_order_groups(uint64_t n, const uint64_t* colptr, const uint64_t* rowval, const int32_t* group,
                      int32_t ngroups, int ordering, uint64_t* perm_out) {
  std::vector<int64_t> Ap(n + 1);
  std::vector<int32_t> Ai(colptr[n]);
  for (uint64_t j = 0; j <= n; j++) Ap[j] = (int64_t)colptr[j];
  for (uint64_t p = 0; p < colptr[n]; p++) Ai[p] = (int32_t)rowval[p];
  cb::SymbolicOptions so;
  so.ordering = ordering ? ordering : cb::ORDER_BEST;
  std::vector<int> perm;
  int kind = 0;
  int rc = cb::order_with_groups((int)n, Ap.data(), Ai.data(), group, ngroups, so, perm, &kind);
  if (rc) return rc;
  for (uint64_t k = 0; k < n; k++) perm_out[k] = (uint64_t)perm[k];
  return kind;
}

// stand-alone orderings (perm_out length n)
int csym_order(uint64_t n, const uint64_t* colptr, const uint64_t* rowval, int kind,
               double dense_scale, int nd_leaf, uint64_t* perm_out) {
  std::vector<int64_t> Ap(n + 1);
  uint64_t nnz = colptr[n];
  std::vector<int32_t> Ai(nnz);
  for (uint64_t j = 0; j <= n; j++) Ap[j] = (int64_t)colptr[j];
  for (uint64_t p = 0; p < nnz; p++) Ai[p] = (int32_t)rowval[p];
  std::vector<int> perm;
  if (kind == 1) cb::amd_order((int)n, Ap.data(), Ai.data(), dense_scale, perm);
  else cb::nd_order((int)n, Ap.data(), Ai.data(), dense_scale, nd_leaf > 0 ? nd_leaf : 200, perm);
  if (perm.size() != n) return -1;
  for (uint64_t k = 0; k < n; k++) perm_out[k] = (uint64_t)perm[k];
  return 0;
}

// subtree-to-rank mapping of one factorisation (symbolic.h: ShardPlan).  owner_out[nsup]; stats_out[8] =
// {total_flops, top_flops, max rank flops, min rank flops, exchange_doubles, exchange_vec, top_levels, model_speedup}
int csym_shard_plan(int64_t nsup, const int64_t* sn_first, const int64_t* sn_rowptr, const int64_t* sn_parent,
                    int nranks, int64_t* owner_out, double* stats_out) {
  std::vector<int> f(sn_first, sn_first + nsup + 1), par(sn_parent, sn_parent + nsup);
  cb::ShardPlan P;
  int rc = cb::plan_shards((int)nsup, f.data(), sn_rowptr, par.data(), nranks, P);
  if (rc) return rc;
  for (int64_t s = 0; s < nsup; s++) owner_out[s] = P.owner[s];
  double mx = 0.0, mn = 1e300;
  for (double v : P.rank_flops) { if (v > mx) mx = v; if (v < mn) mn = v; }
  stats_out[0] = P.total_flops; stats_out[1] = P.top_flops; stats_out[2] = mx; stats_out[3] = mn;
  stats_out[4] = (double)P.exchange_doubles; stats_out[5] = (double)P.exchange_vec; stats_out[6] = P.top_levels;
  stats_out[7] = P.model_speedup;
  return 0;
}

}  // extern "C"
