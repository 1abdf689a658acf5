// Exponential and 3-D power cone kernels (sm_100a): one thread per cone.  See cones_nonsym.cuh for the per-cone
// arithmetic and the reference map, cones.h for the cone engine these methods belong to.
//
// Layout: cone k of the nns nonsymmetric cones owns rows off[id]..off[id]+2 of the m-vectors and entries
// boff[id]..boff[id]+5 of the flat Hs vector (packed upper triangle, the KKT block order of
// kkt_assembly.rs:107-183).  Per-cone state (dual Hessian, Hs, dual gradient, scaling-point z) is kept in
// structure-of-arrays form, component j of cone k at [j*nns + k]: thread k and thread k+1 touch adjacent doubles.
//
// The composite step length (compositecone.rs:289-332) threads one running alpha through the cones in order;
// every cone shortens it by whole multiples of the backtracking factor, so the result is alpha0 * step^J with
// J the largest per-cone count: cones count independently, an integer atomicMax combines them (order
// independent, hence reproducible), and one thread rebuilds alpha with the same sequence of multiplications.
#include "cones.h"
#include "cones_nonsym.cuh"

#include <cstdio>

namespace cb {

using ns3::Sym3;

namespace {

__host__ __device__ inline ns3::View view(const ConeDev& c) {
  return ns3::View{c.nns, c.ns_list, c.type, c.off, c.boff, c.ns_alpha, c.ns_Hd, c.ns_Hs, c.ns_grad, c.ns_z};
}
#define NS_THREAD                                            \
  const int k = blockIdx.x * blockDim.x + threadIdx.x;       \
  if (k >= c.nns) return;                                    \
  const ns3::View v = view(c);

__global__ void k_ns_unit_init(ConeDev c, double* __restrict__ z, double* __restrict__ s) {
  NS_THREAD
  ns3::body_unit_init(v, k, z, s);
}
__global__ void k_ns_update_scaling(ConeDev c, const double* __restrict__ s, const double* __restrict__ z, double mu,
                                    int strategy) {
  NS_THREAD
  ns3::body_update_scaling(v, k, s, z, mu, strategy);
}
__global__ void k_ns_get_Hs(ConeDev c, double* __restrict__ Hs, double sign) {
  NS_THREAD
  ns3::body_get_Hs(v, k, Hs, sign);
}
__global__ void k_ns_mul_Hs(ConeDev c, double* __restrict__ y, const double* __restrict__ x) {
  NS_THREAD
  ns3::body_mul_Hs(v, k, y, x);
}
__global__ void k_ns_copy_rows(ConeDev c, double* __restrict__ out, const double* __restrict__ in) {
  NS_THREAD
  ns3::body_copy_rows(v, k, out, in);
}
__global__ void k_ns_combined_shift(ConeDev c, double* __restrict__ shift, const double* __restrict__ step_z,
                                    const double* __restrict__ step_s, double sigmamu) {
  NS_THREAD
  ns3::body_combined_shift(v, k, shift, step_z, step_s, sigmamu);
}
// *alpha is the step the symmetric cones allow; the largest backtracking count goes to *jmax
__global__ void k_ns_step_count(ConeDev c, const double* __restrict__ dz, const double* __restrict__ ds,
                                const double* __restrict__ z, const double* __restrict__ s,
                                const double* __restrict__ alpha, double a_min, double step, int* jmax) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  int j = 0;
  if (k < c.nns) j = ns3::body_step_count(view(c), k, dz, ds, z, s, *alpha, a_min, step);
  __syncwarp();
  for (int o = 16; o > 0; o >>= 1) { const int t = __shfl_xor_sync(0xffffffffu, j, o); j = t > j ? t : j; }
  if ((threadIdx.x & 31) == 0 && j > 0) atomicMax(jmax, j);
}
__global__ void k_ns_step_final(double* alpha, int* jmax, double step) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    *alpha = ns3::body_step_final(*alpha, *jmax, step);
    *jmax = 0;
  }
}

}  // namespace

// ------------------------------------------------------------------------------------------------- host side
#define NCK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { std::fprintf(stderr, "[clarabel_b200] CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return -20; } } while (0)
#define NS_GRID ((dev.nns + 127) / 128)

int ConeSet::ns_prepare(const std::vector<double>& alpha_per_cone) {
  dev.nns = (int)ns_list.size();
  if (dev.nns == 0) return 0;
  std::vector<double> al(ns_list.size());
  for (size_t k = 0; k < ns_list.size(); k++) al[k] = alpha_per_cone[ns_list[k]];
  int* l = nullptr; double* a = nullptr;
  NCK(cudaMalloc((void**)&l, ns_list.size() * sizeof(int)));
  NCK(cudaMemcpy(l, ns_list.data(), ns_list.size() * sizeof(int), cudaMemcpyHostToDevice));
  NCK(cudaMalloc((void**)&a, al.size() * 8));
  NCK(cudaMemcpy(a, al.data(), al.size() * 8, cudaMemcpyHostToDevice));
  dev.ns_list = l; dev.ns_alpha = a;
  const size_t n = ns_list.size();
  NCK(cudaMalloc((void**)&dev.ns_Hd, 6 * n * 8)); NCK(cudaMalloc((void**)&dev.ns_Hs, 6 * n * 8));
  NCK(cudaMalloc((void**)&dev.ns_grad, 3 * n * 8)); NCK(cudaMalloc((void**)&dev.ns_z, 3 * n * 8));
  NCK(cudaMemset(dev.ns_Hd, 0, 6 * n * 8)); NCK(cudaMemset(dev.ns_Hs, 0, 6 * n * 8));
  NCK(cudaMemset(dev.ns_grad, 0, 3 * n * 8)); NCK(cudaMemset(dev.ns_z, 0, 3 * n * 8));
  NCK(cudaMalloc((void**)&dev.ns_jmax, sizeof(int)));
  NCK(cudaMemset(dev.ns_jmax, 0, sizeof(int)));
  return 0;
}
void ConeSet::ns_release() {
  auto fr = [](const void* p) { if (p) cudaFree((void*)p); };
  fr(dev.ns_list); fr(dev.ns_alpha); fr(dev.ns_Hd); fr(dev.ns_Hs); fr(dev.ns_grad); fr(dev.ns_z); fr(dev.ns_jmax);
}

// Cone::unit_initialization of the whole composite cone (compositecone.rs:217-221): zero everything, add the
// unit element of every symmetric cone (e for NN / SOC / PSD, nothing for the zero cone), set the nonsymmetric ones
void ConeSet::unit_initialization(double* z, double* s) {
  if (m == 0) return;
  cudaMemsetAsync(z, 0, (size_t)m * 8, stream);
  cudaMemsetAsync(s, 0, (size_t)m * 8, stream);
  scaled_unit_shift(s, 1.0, true);
  scaled_unit_shift(z, 1.0, false);
  if (dev.nns) { g_launches++; k_ns_unit_init<<<NS_GRID, 128, 0, stream>>>(dev, z, s); }
}
void ConeSet::ns_update_scaling(const double* s, const double* z, double mu, int strategy) {
  if (!dev.nns) return;
  g_launches++;
  k_ns_update_scaling<<<NS_GRID, 128, 0, stream>>>(dev, s, z, mu, strategy);
}
void ConeSet::ns_get_Hs(double* Hs, double sign) {
  if (!dev.nns) return;
  g_launches++;
  k_ns_get_Hs<<<NS_GRID, 128, 0, stream>>>(dev, Hs, sign);
}
void ConeSet::ns_mul_Hs(double* y, const double* x) {
  if (!dev.nns) return;
  g_launches++;
  k_ns_mul_Hs<<<NS_GRID, 128, 0, stream>>>(dev, y, x);
}
void ConeSet::ns_copy_rows(double* out, const double* in) {
  if (!dev.nns) return;
  g_launches++;
  k_ns_copy_rows<<<NS_GRID, 128, 0, stream>>>(dev, out, in);
}
void ConeSet::ns_combined_shift(double* shift, const double* step_z, const double* step_s, double sigmamu) {
  if (!dev.nns) return;
  g_launches++;
  k_ns_combined_shift<<<NS_GRID, 128, 0, stream>>>(dev, shift, step_z, step_s, sigmamu);
}
void ConeSet::ns_step_length(const double* dz, const double* ds, const double* z, const double* s, double* alpha_slot) {
  if (!dev.nns) return;
  g_launches += 2;
  k_ns_step_count<<<NS_GRID, 128, 0, stream>>>(dev, dz, ds, z, s, alpha_slot, ns_amin, ns_step, dev.ns_jmax);
  k_ns_step_final<<<1, 32, 0, stream>>>(alpha_slot, dev.ns_jmax, ns_step);
}

// Cone::compute_barrier summed over all cones (compositecone.rs:334-345) into out[0]; deterministic two-level
// sums per cone class.  partial = 4 device doubles of scratch.
void ConeSet::compute_barrier(const double* z, const double* s, const double* dz, const double* ds, double alpha,
                              double* partial, double* out) {
  const ConeDev c = dev;
  cudaMemsetAsync(partial, 0, 4 * 8, stream);
  if (m) {
    g_launches++;
    k_sum<<<red_grid(m), RED_THREADS, 0, stream>>>(m, [=] __device__(int i) {
      if (c.rowtag[i] != CT_NONNEG) return 0.0;
      return -ns3::lsafe((s[i] + alpha * ds[i]) * (z[i] + alpha * dz[i]));     // nonnegativecone.rs:155-166
    }, ws, partial + 0);
  }
  if (c.nsoc) {
    g_launches++;
    k_sum<<<red_grid(c.nsoc), RED_THREADS, 0, stream>>>(c.nsoc, [=] __device__(int k) {
      const int id = c.soc_list[k], o = c.off[id];
      return ns3::soc_barrier(z + o, s + o, dz + o, ds + o, c.dim[id], alpha);
    }, ws, partial + 1);
  }
  if (c.npsd) {
    g_launches++;
    k_sum<<<red_grid(c.npsd), RED_THREADS, 0, stream>>>(c.npsd, [=] __device__(int k) {
      const int id = c.psd_list[k], o = c.off[id], n = c.psd_n[id];
      double W[CB_PSD_MAX_N * CB_PSD_MAX_N];
      return ns3::psd_neg_logdet(z + o, dz + o, n, alpha, W) + ns3::psd_neg_logdet(s + o, ds + o, n, alpha, W);
    }, ws, partial + 2);
  }
  if (c.nns) {
    g_launches++;
    const ns3::View v = view(c);
    k_sum<<<red_grid(c.nns), RED_THREADS, 0, stream>>>(c.nns, [=] __device__(int k) {
      return ns3::body_barrier(v, k, z, s, dz, ds, alpha);
    }, ws, partial + 3);
  }
  g_launches++;
  k_map<<<1, 32, 0, stream>>>(1, [=] __device__(int) { out[0] = ((partial[0] + partial[1]) + partial[2]) + partial[3]; });
}

}  // namespace cb
