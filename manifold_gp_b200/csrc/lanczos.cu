// Lanczos support kernels: full re-orthogonalisation as a fused tall-skinny GEMV pair, plus small vector helpers.
//
// Algorithm: linear_operator.utils.lanczos.lanczos_tridiag (third party; restated in oracle/solvers.py; reference call
// site operators/graph_laplacian_operator.py:132-135).  linear_operator forms `r.unsqueeze(0).mul(q_mat[:k+1]).sum(..)`
// and `q_mat[:k+1].mul(correction).sum(0)`, materialising two [k+1, N] temporaries per pass.  Here
//     c = Q[0:j]^T r     (lanczos_dots: every Q vector streamed once, deterministic two-stage reduction)
//     r -= Q[0:j] c      (lanczos_axpy_many: every Q vector streamed once), fused with |r|^2
// Q is vector-major: Q[j] is a contiguous length-n vector, so both passes are fully coalesced and HBM bound
// (2 x j x n x w bytes per re-orthogonalisation pass).
#include "common.cuh"

namespace mgp {

constexpr int kLzBlock = 256;
constexpr int kLzRowsPerThread = 4;
constexpr int kLzRowsPerBlock = kLzBlock * kLzRowsPerThread;  // 1024
constexpr int kLzJB = 16;                                      // Q vectors per block in the dots pass

// partials[rb * j + jj] = sum over the row block rb of Q[jj][i] * r[i]
template <typename T>
__global__ void __launch_bounds__(kLzBlock)
lz_dots_kernel(const T* __restrict__ q, int64_t ldq, int j, const T* __restrict__ r, int64_t n, T* __restrict__ partials) {
  __shared__ T sm[kLzBlock / 32][kLzJB];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t row0 = (int64_t)blockIdx.x * kLzRowsPerBlock;
  const int j0 = blockIdx.y * kLzJB;
  T rv[kLzRowsPerThread];
#pragma unroll
  for (int t = 0; t < kLzRowsPerThread; ++t) {
    const int64_t i = row0 + tid + (int64_t)t * kLzBlock;
    rv[t] = i < n ? r[i] : T(0);
  }
#pragma unroll 4
  for (int jj = 0; jj < kLzJB; ++jj) {
    T s = T(0);
    if (j0 + jj < j) {
      const T* qv = q + (int64_t)(j0 + jj) * ldq;
#pragma unroll
      for (int t = 0; t < kLzRowsPerThread; ++t) {
        const int64_t i = row0 + tid + (int64_t)t * kLzBlock;
        if (i < n) s = fma(ld_stream(qv + i), rv[t], s);
      }
    }
    s = warp_sum(s);
    if (lane == 0) sm[warp][jj] = s;
  }
  __syncthreads();
  if (tid < kLzJB && j0 + tid < j) {
    T s = T(0);
#pragma unroll
    for (int w = 0; w < kLzBlock / 32; ++w) s += sm[w][tid];
    partials[(int64_t)blockIdx.x * j + j0 + tid] = s;
  }
}

// c[jj] = sum_rb partials[rb*j + jj]   (fixed order)
template <typename T>
__global__ void lz_dots_reduce_kernel(const T* __restrict__ partials, int nrb, int j, T* __restrict__ c) {
  const int jj = blockIdx.x * blockDim.x + threadIdx.x;
  if (jj >= j) return;
  T s = T(0);
  for (int rb = 0; rb < nrb; ++rb) s += partials[(int64_t)rb * j + jj];
  c[jj] = s;
}

// r[i] -= sum_jj c[jj] * Q[jj][i]; partial |r|^2 per block -> nrm2 via last-block reduce
template <typename T>
__global__ void __launch_bounds__(kLzBlock)
lz_axpy_many_kernel(const T* __restrict__ q, int64_t ldq, int j, T* __restrict__ r, int64_t n, const T* __restrict__ c,
                    T* __restrict__ nrm2, T* __restrict__ partials, unsigned int* counter) {
  extern __shared__ unsigned char smem_raw[];
  T* cs = reinterpret_cast<T*>(smem_raw);  // [j]
  for (int jj = threadIdx.x; jj < j; jj += blockDim.x) cs[jj] = c[jj];
  __syncthreads();
  T local = T(0);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    T acc0 = T(0), acc1 = T(0), acc2 = T(0), acc3 = T(0);
    int jj = 0;
    for (; jj + 4 <= j; jj += 4) {
      acc0 = fma(cs[jj + 0], ld_stream(q + (int64_t)(jj + 0) * ldq + i), acc0);
      acc1 = fma(cs[jj + 1], ld_stream(q + (int64_t)(jj + 1) * ldq + i), acc1);
      acc2 = fma(cs[jj + 2], ld_stream(q + (int64_t)(jj + 2) * ldq + i), acc2);
      acc3 = fma(cs[jj + 3], ld_stream(q + (int64_t)(jj + 3) * ldq + i), acc3);
    }
    for (; jj < j; ++jj) acc0 = fma(cs[jj], ld_stream(q + (int64_t)jj * ldq + i), acc0);
    const T rn = r[i] - ((acc0 + acc1) + (acc2 + acc3));
    r[i] = rn;
    local = fma(rn, rn, local);
  }
  __shared__ T red[kLzBlock / 32];
  local = warp_sum(local);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    T s = T(0);
    for (int w = 0; w < kLzBlock / 32; ++w) s += red[w];
    partials[blockIdx.x] = s;
  }
  if (last_block_ticket(counter)) {
    if (threadIdx.x == 0) {
      T s = T(0);
      for (int b = 0; b < (int)gridDim.x; ++b) s += __ldcg(partials + b);
      *nrm2 = s;
    }
  }
}

// q_out = r / sqrt(nrm2), beta_out = sqrt(nrm2)
template <typename T>
__global__ void lz_normalize_kernel(const T* __restrict__ r, int64_t n, const T* __restrict__ nrm2, T* __restrict__ q_out,
                                    T* __restrict__ beta_out) {
  const T nr = dev_sqrt<T>(*nrm2);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    q_out[i] = r[i] / nr;
  if (beta_out && blockIdx.x == 0 && threadIdx.x == 0) *beta_out = nr;
}

struct LzWs {
  unsigned int* counter;
  void* partials_axpy;   // [grid] T
  void* partials_dots;   // [nrb * j] T
};

static inline int lz_nrb(int64_t n) { return (int)ceil_div(n, kLzRowsPerBlock); }
static inline int lz_axpy_grid(int64_t n) {
  int64_t g = ceil_div(n, kLzBlock);
  const int64_t cap = (int64_t)kNumSMs * 8;
  return (int)(g > cap ? cap : (g < 1 ? 1 : g));
}

static inline LzWs lz_layout(void* ws) {
  LzWs w;
  char* b = (char*)ws;
  w.counter = (unsigned int*)b;
  w.partials_axpy = b + 256;
  w.partials_dots = b + 256 + (size_t)kNumSMs * 8 * 8;
  return w;
}

template <typename T>
static int lz_dots(const T* q, int64_t ldq, int j, const T* r, int64_t n, T* c, void* ws, cudaStream_t st) {
  MGP_CHECK_ARG(q && r && c && ws && j > 0 && n > 0 && ldq >= n, "lanczos_dots: bad arguments");
  LzWs w = lz_layout(ws);
  const int nrb = lz_nrb(n);
  dim3 grid(nrb, (unsigned)ceil_div(j, kLzJB));
  lz_dots_kernel<T><<<grid, kLzBlock, 0, st>>>(q, ldq, j, r, n, (T*)w.partials_dots);
  MGP_LAUNCH_CHECK();
  lz_dots_reduce_kernel<T><<<(unsigned)ceil_div(j, 128), 128, 0, st>>>((const T*)w.partials_dots, nrb, j, c);
  MGP_LAUNCH_CHECK();
  return MGP_OK;
}

template <typename T>
static int lz_axpy(const T* q, int64_t ldq, int j, T* r, int64_t n, const T* c, T* nrm2, void* ws, cudaStream_t st);

template <typename T>
static int lz_reorth(const T* q, int64_t ldq, int j, T* r, int64_t n, T* c, T* nrm2, void* ws, cudaStream_t st) {
  MGP_CHECK_ARG(nrm2 != nullptr && r && ws && n > 0 && j >= 0, "lanczos_reorth: bad arguments");
  if (j > 0) {
    int rc = lz_dots<T>(q, ldq, j, r, n, c, ws, st);
    if (rc) return rc;
  }
  return lz_axpy<T>(q, ldq, j, r, n, c, nrm2, ws, st);
}

template <typename T>
static int lz_axpy(const T* q, int64_t ldq, int j, T* r, int64_t n, const T* c, T* nrm2, void* ws, cudaStream_t st) {
  MGP_CHECK_ARG(nrm2 != nullptr && r && ws && n > 0 && j >= 0 && (j == 0 || (q && c)), "lanczos_axpy: bad arguments");
  LzWs w = lz_layout(ws);
  const size_t smem = (size_t)j * sizeof(T);
  MGP_CHECK_ARG(smem <= 200 * 1024, "lanczos_reorth: j = %d too large for the shared-memory coefficient cache", j);
  if (smem > 48 * 1024)
    MGP_CUDA(cudaFuncSetAttribute(lz_axpy_many_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  lz_axpy_many_kernel<T><<<lz_axpy_grid(n), kLzBlock, smem, st>>>(q, ldq, j, r, n, c, nrm2, (T*)w.partials_axpy, w.counter);
  MGP_LAUNCH_CHECK();
  return MGP_OK;
}

}  // namespace mgp

using namespace mgp;

extern "C" {

size_t mgp_lanczos_ws_bytes(int64_t n, int32_t j) {
  return 256 + (size_t)kNumSMs * 8 * 8 + (size_t)lz_nrb(n) * (size_t)(j > 0 ? j : 1) * 8;
}

int mgp_lanczos_reorth_f32(const float* q, int64_t ldq, int32_t j, float* r, int64_t n, float* c, float* nrm2, void* ws,
                           void* stream) {
  return lz_reorth<float>(q, ldq, j, r, n, c, nrm2, ws, (cudaStream_t)stream);
}
int mgp_lanczos_reorth_f64(const double* q, int64_t ldq, int32_t j, double* r, int64_t n, double* c, double* nrm2,
                           void* ws, void* stream) {
  return lz_reorth<double>(q, ldq, j, r, n, c, nrm2, ws, (cudaStream_t)stream);
}
int mgp_lanczos_axpy_f32(const float* q, int64_t ldq, int32_t j, float* r, int64_t n, const float* c, float* nrm2, void* ws,
                         void* stream) {
  return lz_axpy<float>(q, ldq, j, r, n, c, nrm2, ws, (cudaStream_t)stream);
}
int mgp_lanczos_axpy_f64(const double* q, int64_t ldq, int32_t j, double* r, int64_t n, const double* c, double* nrm2,
                         void* ws, void* stream) {
  return lz_axpy<double>(q, ldq, j, r, n, c, nrm2, ws, (cudaStream_t)stream);
}
int mgp_lanczos_normalize_f32(const float* r, int64_t n, const float* nrm2, float* q_out, float* beta_out, void* stream) {
  MGP_CHECK_ARG(r && nrm2 && q_out && n > 0, "lanczos_normalize: bad arguments");
  lz_normalize_kernel<float><<<lz_axpy_grid(n), kLzBlock, 0, (cudaStream_t)stream>>>(r, n, nrm2, q_out, beta_out);
  MGP_LAUNCH_CHECK();
  return MGP_OK;
}
int mgp_lanczos_normalize_f64(const double* r, int64_t n, const double* nrm2, double* q_out, double* beta_out, void* stream) {
  MGP_CHECK_ARG(r && nrm2 && q_out && n > 0, "lanczos_normalize: bad arguments");
  lz_normalize_kernel<double><<<lz_axpy_grid(n), kLzBlock, 0, (cudaStream_t)stream>>>(r, n, nrm2, q_out, beta_out);
  MGP_LAUNCH_CHECK();
  return MGP_OK;
}
int mgp_lanczos_dots_f32(const float* q, int64_t ldq, int32_t j, const float* r, int64_t n, float* c, void* ws,
                         void* stream) {
  return lz_dots<float>(q, ldq, j, r, n, c, ws, (cudaStream_t)stream);
}
int mgp_lanczos_dots_f64(const double* q, int64_t ldq, int32_t j, const double* r, int64_t n, double* c, void* ws,
                         void* stream) {
  return lz_dots<double>(q, ldq, j, r, n, c, ws, (cudaStream_t)stream);
}

}  // extern "C"
