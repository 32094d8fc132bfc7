// Laplacian value build for one graph bandwidth eps (once per optimiser step; the structure is reused).
//
// Reference arithmetic: manifold_gp/operators/graph_laplacian_operator.py:52-106 (cached properties
// adjacency_unnorm_mat, degree_unnorm_mat, adjacency_mat, degree_mat, laplacian_diag, laplacian_triu).
// The reference accumulates the two degree vectors with four scatter_add_ passes (atomics on CUDA, so the
// summation order changes run to run); here every row sum is a deterministic sub-warp reduction over the CSR row.
//
// Three streaming passes over (col, d2) -- HBM bound, 2 x nnz x (4 + w) + nnz x w bytes in total:
//   pass 1: Dt_i = [1] + sum_p W_p                      W_p = exp(d2_p / (-4 eps^2))
//   pass 2: D_i  = [Dt_i^-2] + sum_p W_p / (Dt_i Dt_j)
//   pass 3: a_p  = W_p / (Dt_i Dt_j) / (sqrt(D_i) sqrt(D_j)) / eps^2 ;  diag_i
#include "common.cuh"

namespace mgp {

constexpr int kValLanes = 8;  // lanes per row (rows hold k-1 .. ~1.6(k-1) entries)

template <typename T, int PASS>
__global__ void __launch_bounds__(256)
lap_values_kernel(const int* __restrict__ rowptr, const int* __restrict__ col, const T* __restrict__ d2,
                  int64_t n, const T* __restrict__ eps_p, int self_loops, T* __restrict__ dt, T* __restrict__ dg,
                  T* __restrict__ diag, T* __restrict__ a) {
  const int lane = threadIdx.x & 31;
  const int l = lane & (kValLanes - 1);
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / kValLanes;
  const T eps = *eps_p;
  const T eps2 = eps * eps;
  const T m4e2 = T(-4) * eps2;
  int p0 = 0, p1 = 0;
  if (row < n) {
    p0 = rowptr[row];
    p1 = rowptr[row + 1];
  }
  if (PASS == 1) {
    T s = T(0);
    for (int p = p0 + l; p < p1; p += kValLanes) s += dev_exp<T>(ld_stream(d2 + p) / m4e2);
    s = subwarp_sum(s, 1, kValLanes);
    if (l == 0 && row < n) dt[row] = (self_loops ? T(1) : T(0)) + s;
  } else if (PASS == 2) {
    const T dti = row < n ? dt[row] : T(1);
    T s = T(0);
    for (int p = p0 + l; p < p1; p += kValLanes) {
      const T w = dev_exp<T>(ld_stream(d2 + p) / m4e2);
      const T dtj = dt[ld_stream(col + p)];
      s += w / (dti * dtj);
    }
    s = subwarp_sum(s, 1, kValLanes);
    if (l == 0 && row < n) dg[row] = (self_loops ? T(1) / (dti * dti) : T(0)) + s;
  } else {
    const T dti = row < n ? dt[row] : T(1);
    const T dgi = row < n ? dg[row] : T(1);
    const T sdi = dev_sqrt<T>(dgi);
    for (int p = p0 + l; p < p1; p += kValLanes) {
      const T w = dev_exp<T>(ld_stream(d2 + p) / m4e2);
      const int j = ld_stream(col + p);
      const T at = w / (dti * dt[j]);
      a[p] = at / (sdi * dev_sqrt<T>(dg[j])) / eps2;
    }
    if (l == 0 && row < n) {
      diag[row] = self_loops ? (T(1) - (T(1) / (dti * dti)) * (T(1) / dgi)) / eps2 : T(1) / eps2;
    }
  }
}

// ---- backward of the value build w.r.t. eps (forward-mode tangents, then one fused reduction) -------------------
//   W' = W d2/(2 eps^3);  t_i = Dt'_i/Dt_i;  At' = At (d2/(2 eps^3) - t_i - t_j);  u_i = D'_i/D_i with
//   D'_i = -2 [Dt_i^-2] t_i + sum_j At'_ij;  a' = a (d2/(2 eps^3) - t_i - t_j - u_i/2 - u_j/2 - 2/eps);
//   diag'_i = -2 diag_i/eps + [ (Dt_i^-2/D_i) (2 t_i + u_i) / eps^2 ]
//   g_eps = sum_p g_a[p] a'[p] + sum_i ( g_diag_i diag'_i + g_dt_i Dt'_i + g_dg_i D'_i )
template <typename T, int PASS>
__global__ void __launch_bounds__(256)
lap_values_grad_kernel(const int* __restrict__ rowptr, const int* __restrict__ col, const T* __restrict__ d2, int64_t n,
                       const T* __restrict__ eps_p, int self_loops, const T* __restrict__ dt, const T* __restrict__ dg,
                       const T* __restrict__ diag, const T* __restrict__ a, const T* __restrict__ g_dt,
                       const T* __restrict__ g_dg, const T* __restrict__ g_diag, const T* __restrict__ g_a,
                       T* __restrict__ tvec, T* __restrict__ uvec, T* __restrict__ partials, unsigned int* counter,
                       T* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int l = lane & (kValLanes - 1);
  const T eps = *eps_p;
  const T eps2 = eps * eps;
  const T m4e2 = T(-4) * eps2;
  const T inv2e3 = T(1) / (T(2) * eps2 * eps);
  T local = T(0);
  const int64_t rows_per_grid = (int64_t)gridDim.x * (blockDim.x / kValLanes);
  for (int64_t row0 = (int64_t)blockIdx.x * (blockDim.x / kValLanes); row0 < n; row0 += rows_per_grid) {
    const int64_t row = row0 + threadIdx.x / kValLanes;
    int p0 = 0, p1 = 0;
    if (row < n) { p0 = rowptr[row]; p1 = rowptr[row + 1]; }
    const T dti = row < n ? dt[row] : T(1);
    if (PASS == 1) {
      T s = T(0);
      for (int p = p0 + l; p < p1; p += kValLanes) {
        const T dd = ld_stream(d2 + p);
        s += dev_exp<T>(dd / m4e2) * dd * inv2e3;
      }
      s = subwarp_sum(s, 1, kValLanes);
      if (l == 0 && row < n) tvec[row] = s / dti;
    } else if (PASS == 2) {
      const T ti = row < n ? tvec[row] : T(0);
      T s = T(0);
      for (int p = p0 + l; p < p1; p += kValLanes) {
        const T dd = ld_stream(d2 + p);
        const int j = ld_stream(col + p);
        const T at = dev_exp<T>(dd / m4e2) / (dti * dt[j]);
        s += at * (dd * inv2e3 - ti - tvec[j]);
      }
      s = subwarp_sum(s, 1, kValLanes);
      if (l == 0 && row < n) {
        const T dprime = (self_loops ? T(-2) * ti / (dti * dti) : T(0)) + s;
        uvec[row] = dprime / dg[row];
      }
    } else {
      const T ti = row < n ? tvec[row] : T(0);
      const T ui = row < n ? uvec[row] : T(0);
      for (int p = p0 + l; p < p1; p += kValLanes) {
        const T dd = ld_stream(d2 + p);
        const int j = ld_stream(col + p);
        const T fac = dd * inv2e3 - ti - tvec[j] - T(0.5) * (ui + uvec[j]) - T(2) / eps;
        local += g_a[p] * a[p] * fac;
      }
      if (l == 0 && row < n) {
        const T dgi = dg[row];
        T dprime = T(-2) * diag[row] / eps;
        if (self_loops) dprime += (T(1) / (dti * dti * dgi)) * (T(2) * ti + ui) / eps2;
        local += g_diag[row] * dprime;
        if (g_dt) local += g_dt[row] * ti * dti;
        if (g_dg) local += g_dg[row] * ui * dgi;
      }
    }
  }
  if (PASS == 3) {
    __shared__ T red[256 / 32];
    local = warp_sum(local);
    if (lane == 0) red[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
      T s = T(0);
      for (int w = 0; w < 256 / 32; ++w) s += red[w];
      partials[blockIdx.x] = s;
    }
    if (last_block_ticket(counter)) {
      if (threadIdx.x == 0) {
        T s = T(0);
        for (int b = 0; b < (int)gridDim.x; ++b) s += __ldcg(partials + b);
        *out = s;
      }
    }
  }
}

template <typename T>
static int lap_values_grad(const int* rowptr, const int* col, const T* d2, int64_t n, const T* eps, int self_loops,
                           const T* dt, const T* dg, const T* diag, const T* a, const T* g_dt, const T* g_dg,
                           const T* g_diag, const T* g_a, T* out, void* ws, cudaStream_t st) {
  MGP_CHECK_ARG(rowptr && col && d2 && eps && dt && dg && diag && a && g_diag && g_a && out && ws, "lap_values_grad: null pointer");
  MGP_CHECK_ARG(n > 0, "lap_values_grad: n must be positive");
  char* b = (char*)ws;
  unsigned int* counter = (unsigned int*)b;
  T* partials = (T*)(b + 256);
  T* tvec = (T*)(b + 256 + (size_t)kNumSMs * 8 * 8);
  T* uvec = tvec + n;
  int64_t grid = ceil_div(n * kValLanes, 256);
  if (grid > (int64_t)kNumSMs * 8) grid = (int64_t)kNumSMs * 8;
  lap_values_grad_kernel<T, 1><<<(unsigned)grid, 256, 0, st>>>(rowptr, col, d2, n, eps, self_loops, dt, dg, diag, a, g_dt, g_dg, g_diag, g_a, tvec, uvec, partials, counter, out);
  MGP_LAUNCH_CHECK();
  lap_values_grad_kernel<T, 2><<<(unsigned)grid, 256, 0, st>>>(rowptr, col, d2, n, eps, self_loops, dt, dg, diag, a, g_dt, g_dg, g_diag, g_a, tvec, uvec, partials, counter, out);
  MGP_LAUNCH_CHECK();
  lap_values_grad_kernel<T, 3><<<(unsigned)grid, 256, 0, st>>>(rowptr, col, d2, n, eps, self_loops, dt, dg, diag, a, g_dt, g_dg, g_diag, g_a, tvec, uvec, partials, counter, out);
  MGP_LAUNCH_CHECK();
  return MGP_OK;
}

template <typename T>
static int lap_values(const int* rowptr, const int* col, const T* d2, int64_t n, const T* eps, int self_loops, T* dt,
                      T* dg, T* diag, T* a, cudaStream_t st) {
  MGP_CHECK_ARG(rowptr && col && d2 && eps && dt && dg && diag && a, "lap_values: null pointer");
  MGP_CHECK_ARG(n > 0, "lap_values: n must be positive");
  const int block = 256;
  const int64_t grid = ceil_div(n * kValLanes, block);
  MGP_CHECK_ARG(grid < ((int64_t)1 << 31), "lap_values: n too large for one launch");
  lap_values_kernel<T, 1><<<(unsigned)grid, block, 0, st>>>(rowptr, col, d2, n, eps, self_loops, dt, dg, diag, a);
  MGP_LAUNCH_CHECK();
  lap_values_kernel<T, 2><<<(unsigned)grid, block, 0, st>>>(rowptr, col, d2, n, eps, self_loops, dt, dg, diag, a);
  MGP_LAUNCH_CHECK();
  lap_values_kernel<T, 3><<<(unsigned)grid, block, 0, st>>>(rowptr, col, d2, n, eps, self_loops, dt, dg, diag, a);
  MGP_LAUNCH_CHECK();
  return MGP_OK;
}

// One pass of the value build on the rows of a row-partitioned structure (SURVEY.md 8e "value build: two halo gathers"):
// n = rows this rank owns; col indexes the rank's extended numbering [own rows | halo rows]; dt / dg hold n_ext entries, the
// halo part of dt must be filled (halo exchange) before pass 2, the halo part of dg before pass 3.
template <typename T>
static int lap_values_pass(int pass, const int* rowptr, const int* col, const T* d2, int64_t n, const T* eps, int self_loops,
                           T* dt, T* dg, T* diag, T* a, cudaStream_t st) {
  MGP_CHECK_ARG(rowptr && col && d2 && eps && dt, "lap_values_pass: null pointer");
  MGP_CHECK_ARG(pass >= 1 && pass <= 3 && (pass < 2 || dg) && (pass < 3 || (diag && a)), "lap_values_pass: pass %d lacks its outputs", pass);
  MGP_CHECK_ARG(n > 0, "lap_values_pass: n must be positive");
  const int block = 256;
  const int64_t grid = ceil_div(n * kValLanes, block);
  MGP_CHECK_ARG(grid < ((int64_t)1 << 31), "lap_values_pass: n too large for one launch");
  if (pass == 1) lap_values_kernel<T, 1><<<(unsigned)grid, block, 0, st>>>(rowptr, col, d2, n, eps, self_loops, dt, dg, diag, a);
  else if (pass == 2) lap_values_kernel<T, 2><<<(unsigned)grid, block, 0, st>>>(rowptr, col, d2, n, eps, self_loops, dt, dg, diag, a);
  else lap_values_kernel<T, 3><<<(unsigned)grid, block, 0, st>>>(rowptr, col, d2, n, eps, self_loops, dt, dg, diag, a);
  MGP_LAUNCH_CHECK();
  return MGP_OK;
}

}  // namespace mgp

extern "C" {

int mgp_lap_values_pass_f32(int32_t pass, const int32_t* rowptr, const int32_t* col, const float* d2csr, int64_t n, const float* eps,
                            int32_t self_loops, float* deg_unnorm, float* deg, float* diag, float* a, void* stream) {
  return mgp::lap_values_pass<float>(pass, rowptr, col, d2csr, n, eps, self_loops, deg_unnorm, deg, diag, a, (cudaStream_t)stream);
}
int mgp_lap_values_pass_f64(int32_t pass, const int32_t* rowptr, const int32_t* col, const double* d2csr, int64_t n, const double* eps,
                            int32_t self_loops, double* deg_unnorm, double* deg, double* diag, double* a, void* stream) {
  return mgp::lap_values_pass<double>(pass, rowptr, col, d2csr, n, eps, self_loops, deg_unnorm, deg, diag, a, (cudaStream_t)stream);
}

int mgp_lap_values_f32(const int32_t* rowptr, const int32_t* col, const float* d2csr, int64_t n, const float* eps,
                       int32_t self_loops, float* deg_unnorm, float* deg, float* diag, float* a, void* stream) {
  return mgp::lap_values<float>(rowptr, col, d2csr, n, eps, self_loops, deg_unnorm, deg, diag, a, (cudaStream_t)stream);
}

int mgp_lap_values_f64(const int32_t* rowptr, const int32_t* col, const double* d2csr, int64_t n, const double* eps,
                       int32_t self_loops, double* deg_unnorm, double* deg, double* diag, double* a, void* stream) {
  return mgp::lap_values<double>(rowptr, col, d2csr, n, eps, self_loops, deg_unnorm, deg, diag, a, (cudaStream_t)stream);
}

size_t mgp_lap_values_grad_ws_bytes(int64_t n) { return 256 + (size_t)mgp::kNumSMs * 8 * 8 + (size_t)n * 2 * 8; }

int mgp_lap_values_grad_f32(const int32_t* rowptr, const int32_t* col, const float* d2csr, int64_t n, const float* eps,
                            int32_t self_loops, const float* deg_unnorm, const float* deg, const float* diag,
                            const float* a, const float* g_deg_unnorm, const float* g_deg, const float* g_diag,
                            const float* g_a, float* g_eps, void* ws, void* stream) {
  return mgp::lap_values_grad<float>(rowptr, col, d2csr, n, eps, self_loops, deg_unnorm, deg, diag, a, g_deg_unnorm, g_deg,
                                     g_diag, g_a, g_eps, ws, (cudaStream_t)stream);
}

int mgp_lap_values_grad_f64(const int32_t* rowptr, const int32_t* col, const double* d2csr, int64_t n, const double* eps,
                            int32_t self_loops, const double* deg_unnorm, const double* deg, const double* diag,
                            const double* a, const double* g_deg_unnorm, const double* g_deg, const double* g_diag,
                            const double* g_a, double* g_eps, void* ws, void* stream) {
  return mgp::lap_values_grad<double>(rowptr, col, d2csr, n, eps, self_loops, deg_unnorm, deg, diag, a, g_deg_unnorm, g_deg,
                                      g_diag, g_a, g_eps, ws, (cudaStream_t)stream);
}

}  // extern "C"
