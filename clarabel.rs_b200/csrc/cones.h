// Device cone engine (internal header): Zero / Nonnegative / second-order cones.
//
// Device counterpart of the reference's `Cone` trait and CompositeCone dispatch
// (/root/reference/src/solver/core/cones/mod.rs:42-154, compositecone.rs:197-352).
// The reference loops over cones sequentially on one thread; here every
// operation is one batched launch per cone *class*: an elementwise kernel over
// all rows (row tags select Zero / Nonnegative semantics) plus one CTA per
// second-order cone with block reductions for its dot products and norms.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <vector>

#include "vec.cuh"

namespace cb {

enum { CT_ZERO = 0, CT_NONNEG = 1, CT_SOC = 2, CT_PSD = 3 };
constexpr int SOC_NO_EXPANSION_MAX_SIZE = 4;  // socone.rs:46

struct ConeDev {
  int ncones = 0, m = 0, nsoc = 0;
  const int* type = nullptr;
  const int* off = nullptr;
  const int* dim = nullptr;
  const int* boff = nullptr;     // offset of the cone's Hs block
  const int* sparse = nullptr;   // 1 if sparse-expanded SOC
  const int* soc_list = nullptr; // cone ids of the SOC cones
  const signed char* rowtag = nullptr;  // [m] cone type of every row
  double* w = nullptr;    // [m]
  double* lam = nullptr;  // [m]
  double* eta = nullptr;  // [ncones]
  double* u = nullptr;    // [m] sparse SOC data
  double* v = nullptr;    // [m]
  double* dd = nullptr;   // [ncones]
  int* fail = nullptr;    // scaling failure flag
  // PSD triangle cones (matrix dimension <= CB_PSD_MAX_N)
  int npsd = 0;
  const int* psd_list = nullptr;   // cone ids
  const int* psd_n = nullptr;      // [ncones] matrix dimension (0 for other cones)
  const long long* psd_moff = nullptr;  // [ncones] offset into the n x n matrix arenas
  double *psd_R = nullptr, *psd_Rinv = nullptr, *psd_RRt = nullptr;
};

constexpr int CB_PSD_MAX_N = 32;

// dim = number of rows the cone occupies (numel); psd_n = matrix dimension of a PSD cone
struct ConeSpec { int type; int dim; int psd_n = 0; };

class ConeSet {
 public:
  std::vector<ConeSpec> cones;       // after collapsing
  std::vector<int> off, boff, sparse_flag, soc_list;
  int m = 0, nHs = 0, degree = 0, p = 0;  // p = number of sparse expansion rows
  ConeDev dev;
  cudaStream_t stream = nullptr;
  ReduceWS ws;
  const int* row2blk_dev = nullptr;  // row -> slot in the flat Hs vector (diagonal-block cones)
  double* d_pmin = nullptr;
  double* d_psum = nullptr;

  // collapse like SupportedConeT::new_collapsed (supportedcone.rs:105-161)
  static int collapse(const int32_t* types, const uint64_t* dims, uint64_t n, std::vector<ConeSpec>& out);
  int init(const std::vector<ConeSpec>& cs, cudaStream_t st);
  void release();

  void set_identity_scaling();
  void update_scaling(const double* s, const double* z);  // failure -> dev.fail
  void get_Hs(double* Hs, bool negate);
  void mul_Hs(double* y, const double* x);
  void affine_ds(double* ds);
  void combined_ds_shift(double* shift, double* step_z, double* step_s, double sigmamu);
  void ds_from_dz_offset(double* out, const double* ds, const double* z);
  // alpha slot must be pre-set to alpha_max by the caller (device double)
  void step_length(const double* dz, const double* ds, const double* z, const double* s, double* alpha_slot);
  // out2[0] = min margin, out2[1] = sum of positive margins
  void margins(const double* z, double* out2);
  void scaled_unit_shift(double* z, double alpha, bool primal);

  // PSD pieces (cones_psd.cu)
  std::vector<int> psd_list;
  int psd_nmax = 0, psd_numel_max = 0;
  int psd_prepare();
  void psd_set_identity();
  void psd_update_scaling(const double* s, const double* z);
  void psd_get_Hs(double* Hs, double sign);
  void psd_apply(int op, double* out, double* a, double* b, double scalar);
  void psd_step_length(const double* dz, const double* ds, double* alpha_slot);
  void psd_margins(const double* z, double* pmin, double* psum);
};

}  // namespace cb
