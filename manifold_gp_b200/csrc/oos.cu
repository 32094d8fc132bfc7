// Out-of-sample (Nystrom) extension of eigenvectors to new points: an ELL SpMM with fixed row length k.
//
// Reference: GraphLaplacianOperator.out_of_sample, manifold_gp/operators/graph_laplacian_operator.py:146-157.
// The reference materialises a [Q, k, m] temporary (`out.unsqueeze(-1).mul(x[edge_idx])`, :156); here one warp owns
// one query row: the k weights are normalised in shared memory (two warp reductions), then the lanes sweep the m
// eigenvector columns with coalesced gathers of phi rows.
#include "common.cuh"

namespace mgp {

constexpr int kOosWarps = 8;

template <typename T>
__global__ void __launch_bounds__(kOosWarps * 32)
oos_kernel(const T* __restrict__ d2, const int64_t* __restrict__ idx, int64_t nq, int k, const T* __restrict__ eps_p,
           const T* __restrict__ dt, const T* __restrict__ dg, int normalization, const T* __restrict__ phi,
           int64_t ldphi, int m, T* __restrict__ out, int64_t ldo) {
  extern __shared__ unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  T* w = reinterpret_cast<T*>(smem_raw) + (size_t)warp * k;
  int* id = reinterpret_cast<int*>(reinterpret_cast<T*>(smem_raw) + (size_t)kOosWarps * k) + (size_t)warp * k;
  const T eps = *eps_p;
  const T m4e2 = T(-4) * eps * eps;
  for (int64_t qi = (int64_t)blockIdx.x * kOosWarps + warp; qi < nq; qi += (int64_t)gridDim.x * kOosWarps) {
    T s1 = T(0);
    for (int kk = lane; kk < k; kk += 32) {
      const int j = (int)idx[qi * k + kk];
      const T v = dev_exp<T>(d2[qi * k + kk] / m4e2);   // :147
      id[kk] = j;
      w[kk] = v;
      s1 += v;
    }
    s1 = warp_sum(s1);                                    // degree_test, :148
    T s2 = T(0);
    for (int kk = lane; kk < k; kk += 32) {
      const T v = w[kk] / (dt[id[kk]] * s1);              // :149
      w[kk] = v;
      s2 += v;
    }
    s2 = warp_sum(s2);
    const T rs = normalization == 0 ? dev_sqrt<T>(s2) : s2;
    for (int kk = lane; kk < k; kk += 32) {
      if (normalization == 0) w[kk] = w[kk] / (dev_sqrt<T>(dg[id[kk]]) * rs);   // symmetric, :151-152
      else w[kk] = w[kk] / rs;                                                    // randomwalk, :153-154
    }
    __syncwarp();
    for (int c = lane; c < m; c += 32) {                  // :156
      T acc = T(0);
      for (int kk = 0; kk < k; ++kk) acc = fma(w[kk], __ldg(phi + (int64_t)id[kk] * ldphi + c), acc);
      out[qi * ldo + c] = acc;
    }
    __syncwarp();
  }
}

template <typename T>
static int oos(const T* d2, const int64_t* idx, int64_t nq, int k, const T* eps, const T* dt, const T* dg, int normalization,
               const T* phi, int64_t ldphi, int m, T* out, int64_t ldo, cudaStream_t st) {
  MGP_CHECK_ARG(d2 && idx && eps && dt && dg && phi && out, "out_of_sample: null pointer");
  MGP_CHECK_ARG(nq > 0 && k > 0 && m > 0 && ldphi >= m && ldo >= m, "out_of_sample: bad shape");
  MGP_CHECK_ARG(normalization == 0 || normalization == 1, "out_of_sample: normalization must be 0 (symmetric) or 1 (randomwalk)");
  const size_t smem = (size_t)kOosWarps * k * (sizeof(T) + sizeof(int));
  MGP_CHECK_ARG(smem <= 48 * 1024, "out_of_sample: k = %d too large", k);
  int64_t grid = ceil_div(nq, kOosWarps);
  if (grid > (int64_t)kNumSMs * 8) grid = (int64_t)kNumSMs * 8;
  oos_kernel<T><<<(unsigned)grid, kOosWarps * 32, smem, st>>>(d2, idx, nq, k, eps, dt, dg, normalization, phi, ldphi, m, out, ldo);
  MGP_LAUNCH_CHECK();
  return MGP_OK;
}

}  // namespace mgp

extern "C" {

int mgp_out_of_sample_f32(const float* d2, const int64_t* idx, int64_t nq, int32_t k, const float* eps,
                          const float* deg_unnorm, const float* deg, int32_t normalization, const float* phi,
                          int64_t ldphi, int32_t m, float* out, int64_t ldo, void* stream) {
  return mgp::oos<float>(d2, idx, nq, k, eps, deg_unnorm, deg, normalization, phi, ldphi, m, out, ldo, (cudaStream_t)stream);
}
int mgp_out_of_sample_f64(const double* d2, const int64_t* idx, int64_t nq, int32_t k, const double* eps,
                          const double* deg_unnorm, const double* deg, int32_t normalization, const double* phi,
                          int64_t ldphi, int32_t m, double* out, int64_t ldo, void* stream) {
  return mgp::oos<double>(d2, idx, nq, k, eps, deg_unnorm, deg, normalization, phi, ldphi, m, out, ldo, (cudaStream_t)stream);
}

}  // extern "C"
