// Device-side view and host object of the multifrontal LDL^T (internal header).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <thread>
#include <vector>

#include "../../include/clarabel_b200.h"
#include "symbolic.h"
#include "ldl_solve_plan.h"

#define CB_MAX_PANEL 128
#define CB_PB_MAXNS 64     /* widest panel (columns) of a front */
#define CB_PB_LD 66        /* padded leading dimension of the pivot block in shared memory */
#define CB_SOLVE_SMALL_NS 8 /* solves: fronts with at most this many pivots get one warp, wider ones one CTA */
#define CB_SOLVE_STAGE 1536  /* doubles of gathered x staged in shared memory by k_bwd_big */
#define CB_SMALL_CHILD 16  /* children of big fronts with at most this many rows go through sorted entry lists */
#define CB_BIG_NR 96       /* fronts with at least this many rows below the pivot block use the multi-CTA path */

namespace cb {

enum { ST_REGCOUNT = 0, ST_ZEROPIV = 1, ST_POSINERTIA = 2, ST_NONFINITE = 3, ST_COUNT = 8 };

// Plain-pointer bundle passed by value to kernels.
struct LDLDev {
  const int* sn_first = nullptr;
  const long long* sn_rowptr = nullptr;
  const int* sn_rows = nullptr;
  const long long* child_ptr = nullptr;
  const int* child_list = nullptr;
  const int* rel = nullptr;  // indexed like sn_rows
  const long long* panel_off = nullptr;
  const long long* upd_off = nullptr;
  const long long* asm_ptr = nullptr;
  const int* asm_src = nullptr;
  const long long* asm_dst = nullptr;
  const int* level_tasks = nullptr;
  const int* child_nb = nullptr;      // per front: rows that land inside its parent's pivot block
  const int2* child_trange = nullptr; // per front: [first,last] 64-row tile of the parent's update matrix it touches
  const int* child_tptr_off = nullptr; // per front: offset into child_tptr
  const int* child_tptr = nullptr;     // first own row falling into each parent tile row (tlo..thi+1)
  const signed char* child_small = nullptr;  // 1: child of a big front handled through the sorted entry lists
  const int *sc_panel_ptr = nullptr, *sc_panel_src = nullptr, *sc_panel_dst = nullptr;  // per big front
  const int *sc_tile_ptr = nullptr, *sc_tile_src = nullptr, *sc_tile_dst = nullptr;     // per update tile
  const int* gat_ptr = nullptr;       // solves: per front slot, CSR of contributing child update-vector entries
  const int* gat_src = nullptr;
  const int* perm = nullptr;
  const signed char* dsigns = nullptr;  // permuted order
  double* vals = nullptr;               // KKT values, caller's CSC order
  double* L = nullptr;                  // dense panels
  double* U = nullptr;                  // update-matrix arena
  double* D = nullptr;
  double* Dinv = nullptr;
  double* u = nullptr;                  // per-front update vectors for the solves
  int* status = nullptr;
  double reg_eps = 1e-13, reg_delta = 2e-7;
  int reg_enable = 1;
};

struct LaunchSeg {
  int kind = 0;  // 0 fused small fronts, 1 big-front panels, 2 big-front update tiles
  int level, base, count, smem_doubles, threads;
};

// dataflow factorisation plan (see k_factor_df in ldl.cu)
struct DFFactor {
  int ntask = 0;
  const int4* tasks = nullptr;      // 4 x int4 per task: see DFTask in ldl.cu
  const int* desc = nullptr;        // 12 ints per child record
  int* pend = nullptr;              // [nsup] children still running
  int* diag_done = nullptr;         // [nsup]
  int* rows_left = nullptr;         // [nsup]
  int* tiles_left = nullptr;        // [nsup]
  int* qhead = nullptr;
  const int* parent = nullptr;
  const int* big_pos = nullptr;     // [nsup] index into sc_panel_ptr (position in the big-front list) or -1
  const int* tile_base = nullptr;   // [nsup] first global tile id of the front or -1
  unsigned long long* trace = nullptr;  // optional [ntask][4]: grab, ready, end (globaltimer ns), smid
};


class LDLObject {
 public:
  // triangular solves (ldl_solve.cuh): task queue, counters, level-0 leaf lists, wide fronts whose pivot block is inverted
  SVPlan sv;
  int *d_sv_init = nullptr, *d_sv_cnt = nullptr, *d_sv_wide = nullptr, *d_sv_leaf1 = nullptr, *d_sv_leafn = nullptr, *d_sv_leafw = nullptr;
  size_t sv_ninit = 0, sv_nzero = 0, sv_smem[2] = {0, 0};
  int sv_cap = 0, sv_nwide = 0, sv_nleaf1 = 0, sv_nleafn = 0, sv_nleafw = 0, sv_leafw_nrmax = 0, sv_leafw_grid = 1, sv_ntask_owned = 0;
  std::vector<int> h_sv_tasks;
  std::vector<int> sv_wide_runs;   // (first, widest ns) pairs of the launches of k_invert_pivots, closed by (count, 0)
  int sv_configure();
  int sv_occupancy();
  int sv_reset();
  void sv_sweep(bool fwd, int nrhs, const SVPlan& q, const SVRhs& r);
  void sv_leaves(bool fwd, int nrhs, const SVRhs& r);
  void invert_pivots();
  DFFactor dff;
  int *d_dff_init = nullptr, *d_dff_cnt = nullptr;
  std::vector<int> h_dff_tasks;
  int dff_grid = 0;
  size_t dff_nsup4 = 0;
  bool factor_dataflow = true;
  int df_grid = 0;
  int solve_minb = 4;
  bool use_dataflow = true;
  int n = 0;
  int64_t nnzA = 0;
  int device = 0;
  cldl_opts opts{};
  Symbolic S;
  LDLDev dev;
  std::vector<LaunchSeg> plan;
  std::thread big_alloc;                 // init: allocates the factor storage beside the plan building
  int big_alloc_rc = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t stream_a = nullptr, stream_b = nullptr;   // side streams of the narrow-leaf solve kernels
  cudaEvent_t ev_leaf[3] = {nullptr, nullptr, nullptr};
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  int* h_status = nullptr;
  double* d_xp = nullptr;  // permuted work vector
  double* d_bx = nullptr;  // staging for host-pointer solve: [b ; x]
  int* d_tmp_idx = nullptr;
  double* d_tmp_val = nullptr;
  signed char* d_tmp_sgn = nullptr;
  size_t tmp_cap = 0;
  int* d_big_tasks = nullptr;
  int4* d_tiles = nullptr;
  int64_t n_tiles = 0;
  std::vector<int> h_idx;
  bool factored = false;
  uint64_t regularize_count = 0, positive_inertia = 0;
  // ---- one factorisation on several GPUs: this object is one rank (see ShardPlan in symbolic.h) ----
  int shard_nranks = 1, shard_rank = 0;
  ShardPlan shard;                       // owner[front] = rank or -1 (replicated top)
  std::vector<std::vector<int>> shard_cut;   // per rank: its cut roots (fronts whose parent is in the top part)
  std::vector<std::vector<int>> shard_xidx;  // per rank: caller-order indices of the x entries it computes
  std::vector<int*> d_shard_xidx;            // the same lists on the device
  std::vector<long long*> d_shard_segs[2];   // per what (0 update matrices, 1 update vectors) and rank: (arena offset, buffer offset, length) of its cut roots
  std::vector<int> shard_nsegs[2];
  int shard_seglist(int what, int rank, const long long** d_out, int* nseg);
  int dff_ntask_owned = 0;               // tasks of the owned phase (they come first in the queue)
  int h_phase_start[2] = {0, 0};         // pinned-lifetime host copies of the queue heads the top phases start from
  uint64_t shard_count_owned[2] = {0, 0};    // regularize_count / positive_inertia of the owned phase
  bool sharded() const { return shard_nranks > 1; }
  bool mine(int s) const { return !sharded() || shard.owner[s] == shard_rank || shard.owner[s] < 0; }
  bool owned(int s) const { return !sharded() || shard.owner[s] == shard_rank; }
  int refactor_phase_async(int phase);   // 0: owned subtrees, 1: top part (after the cut roots' update matrices arrived)
  int solve_phase_async(double* d_x, const double* d_b, int phase);   // 0: permute + forward owned, 1: forward top + backward
  // what: 0 update matrices of the cut roots (per refactor), 1 their update vectors (per solve), 2 x entries
  uint64_t shard_count(int what, int rank) const;
  int shard_pack(int what, double* d_buf, const double* d_x);
  int shard_unpack(int what, int rank, const double* d_buf, double* d_x);
  // transport (all-gather between ranks) and the self-driven sharded refactor / solve built on it
  cldl_allgather_fn transport = nullptr;
  void* nccl_comm = nullptr;             // own NCCL communicator: exchanges are stream-ordered (set_nccl)
  unsigned long long n_collectives = 0;  // NCCL all-gathers issued by this handle
  bool has_transport() const { return transport != nullptr || nccl_comm != nullptr; }
  int set_nccl(const char* libpath, const unsigned char* id128, int nranks, int rank);
  void* transport_ctx = nullptr;
  double *d_xsend = nullptr, *d_xrecv = nullptr;
  size_t xbuf_cap = 0;
  int exchange(int what, double* d_x);
  int refactor_sharded();
  int solve_sharded(double* d_x, const double* d_b);

  int init(int n, const int64_t* Ap, const int32_t* Ai, const double* Ax, const int8_t* dsigns,
           const cldl_opts& o, const int* perm_in);
  void release();
  int refactor_async();
  int sync_status();
  // one right-hand side, or two swept together (d_x1 / d_b1 non-null): the panels are read once for both
  int solve_async(double* d_x, const double* d_b, double* d_x1 = nullptr, const double* d_b1 = nullptr);
  double *d_xp2 = nullptr, *d_u2 = nullptr;   // work vectors of the second right-hand side
  int ensure_tmp(size_t len);
  int stage_index(const uint64_t* index, uint64_t len);
};

}  // namespace cb
