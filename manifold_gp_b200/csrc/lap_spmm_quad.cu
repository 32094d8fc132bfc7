// EXPERIMENTAL (not dispatched by default): quad-row SpMM for 64-byte rows of X.  Parity-tested on B200 against the dense
// operator (tests/test_gpu_spmm_kernels.py::test_experimental_quad_row_kernel_matches_dense, fp32 1e-5 / fp64 1e-12, fused dot,
// caller-order translation); NOT yet timed -- the last GPU seconds of round 1 went into that test.
//
//   Y = post .* ( (diag + shift) .* X  -  A X )        16 fp32 / 8 fp64 columns per pass
//
// Same operation and reference lines as lap_spmm_wi.cu (graph_laplacian_operator.py:108-124 / one step of
// precision_matern_operator.py:28-32).  Why: the C = 16 kernel sits on the shared-memory roofline (72.6 % of the pipe, ncu),
// and 64 % of its wavefronts are the 64-byte X-row loads, one per nonzero.  Morton-adjacent rows of the kNN graph share most
// of their columns (union of a row quad = 0.44 of the sum of the four lists, profiles/pair_stats.py), so this kernel walks the
// UNION columns of row quads: one X-row load serves four matrix rows.  Per step a lane group (4 lanes = the four 16-byte
// chunks of a row) reads one tile-local column index, one 4-wide value slot (the quad's four values for that column, zero
// where a row lacks it) and its chunk of the X row, and does 4 x VEC FMAs; the four lane groups of a quad take every 4th
// union column and are summed with two shuffle rounds at the end.  Stream layout: graph.quad_streams (tested on the CPU by
// emulating exactly this walk).  First version: one 512-thread block per tile, X rows staged with cp.async, streams read from
// global memory -- the shape of lap_spmm_tile64_kernel; the pipelined producer of lap_spmm_wi_kernel is the next step.
#include "common.cuh"
#include "pipe_common.cuh"
#include "spmm_common.cuh"

namespace mgp {

constexpr int kQdRows = 128;
constexpr int kQdWarps = 16;
constexpr int kQdThreads = kQdWarps * 32;

template <typename T>
__global__ void __launch_bounds__(kQdThreads)
lap_spmm_quad_kernel(const int* __restrict__ qwptr, const unsigned short* __restrict__ qidx, const T* __restrict__ qval,
                     const int* __restrict__ qrows, const T* __restrict__ diag, const int* __restrict__ hptr,
                     const int* __restrict__ hcol, const T* __restrict__ shift_p, const T* __restrict__ post,
                     const int* __restrict__ xmap, const int* __restrict__ ymap, const T* __restrict__ x, int64_t ldx,
                     T* __restrict__ y, int64_t ldy, int64_t n, int c0, const T* __restrict__ dot_with, T* __restrict__ dot_out,
                     T* __restrict__ partials, unsigned int* __restrict__ counter, int dot_is_x) {
  constexpr int VEC = 16 / sizeof(T);
  constexpr int CW = 4 * VEC;
  constexpr int ROW_BYTES = 64;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char* const xs = smem_raw;
  const int t = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t row0 = (int64_t)t * kQdRows;
  const int nrows = (int)min((int64_t)kQdRows, n - row0);
  const int h0 = hptr[t], nh = hptr[t + 1] - h0;
  const int ch = lane & 3;                               // 16-byte chunk of the 64-byte row this lane owns
  {
    const int c4 = tid & 3;
    const unsigned char* xb = reinterpret_cast<const unsigned char*>(x + c0) + c4 * 16;
    const int64_t ldxb = ldx * (int64_t)sizeof(T);
    for (int rr = tid >> 2; rr < nrows + nh; rr += kQdThreads / 4) {
      int64_t src = rr < nrows ? row0 + rr : (int64_t)__ldg(hcol + h0 + rr - nrows);
      if (xmap) src = __ldg(xmap + src);
      const int dst = rr < nrows ? rr : kQdRows + (rr - nrows);
      cp_async16(xs + (size_t)dst * ROW_BYTES + c4 * 16, xb + src * ldxb);
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
  }
  __syncthreads();
  const T shift = shift_p ? *shift_p : T(0);
  const int g = lane >> 2;                               // lane group 0..7: quad (g >> 2) of the warp, sub-list (g & 3)
  const int base = qwptr[t * kQdWarps + warp];
  const int steps = (qwptr[t * kQdWarps + warp + 1] - base) >> 3;
  const unsigned short* ip = qidx + base + g;
  const T* vp = qval + ((size_t)base + g) * 4;
  T acc[4][VEC];
#pragma unroll
  for (int s = 0; s < 4; ++s)
#pragma unroll
    for (int v = 0; v < VEC; ++v) acc[s][v] = T(0);
#pragma unroll 2
  for (int st = 0; st < steps; ++st) {
    const uint32_t j = __ldcs(ip);
    T q4[4];
    if constexpr (sizeof(T) == 4) {
      const float4 q = __ldcs(reinterpret_cast<const float4*>(vp));
      q4[0] = q.x; q4[1] = q.y; q4[2] = q.z; q4[3] = q.w;
    } else {
      const double2 qa = __ldcs(reinterpret_cast<const double2*>(vp)), qb = __ldcs(reinterpret_cast<const double2*>(vp) + 1);
      q4[0] = qa.x; q4[1] = qa.y; q4[2] = qb.x; q4[3] = qb.y;
    }
    ip += 8;
    vp += 32;
    const Vec<T, VEC> xv = *reinterpret_cast<const Vec<T, VEC>*>(xs + j * ROW_BYTES + ch * 16);
#pragma unroll
    for (int s = 0; s < 4; ++s)
#pragma unroll
      for (int v = 0; v < VEC; ++v) acc[s][v] = fma(q4[s], xv.v[v], acc[s][v]);
  }
  // sum the four sub-lists of a quad: lanes that differ in bits 2 and 3 (same quad, same chunk)
#pragma unroll
  for (int s = 0; s < 4; ++s)
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      acc[s][v] += __shfl_xor_sync(0xffffffffu, acc[s][v], 4);
      acc[s][v] += __shfl_xor_sync(0xffffffffu, acc[s][v], 8);
    }
  // lane group `sub` of a quad finishes row slot `sub`
  const int sub = g & 3;
  const int row = qrows[(((size_t)t * kQdWarps + warp) * 2 + (g >> 2)) * 4 + sub];
  T dsum[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) dsum[v] = T(0);
  if (row >= 0) {
    T res[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) res[v] = sub == 0 ? acc[0][v] : sub == 1 ? acc[1][v] : sub == 2 ? acc[2][v] : acc[3][v];
    const int r = (int)(row - row0);
    const Vec<T, VEC> xi = *reinterpret_cast<const Vec<T, VEC>*>(xs + (size_t)r * ROW_BYTES + ch * 16);
    const T d = __ldg(diag + row) + shift;
    const T po = post ? __ldg(post + row) : T(1);
    Vec<T, VEC> out;
#pragma unroll
    for (int v = 0; v < VEC; ++v) out.v[v] = po * (d * xi.v[v] - res[v]);
    const int64_t yrow = ymap ? (int64_t)__ldg(ymap + row) : (int64_t)row;
    st_vec<T, VEC>(y + yrow * ldy + c0 + ch * VEC, out);
    if (dot_out) {
      Vec<T, VEC> dw = xi;
      if (!dot_is_x) {
        const int64_t drow = xmap ? (int64_t)__ldg(xmap + row) : (int64_t)row;
        dw = ldg_vec<T, VEC>(dot_with + drow * ldx + c0 + ch * VEC);
      }
#pragma unroll
      for (int v = 0; v < VEC; ++v) dsum[v] = fma(dw.v[v], out.v[v], dsum[v]);
    }
  }
  if (dot_out) {
    __syncthreads();
    spmm_dot_epilogue<T, VEC, 4, CW, kQdThreads>(dsum, CW, c0, partials, counter, dot_out);
  }
}

template <typename T>
static int lap_spmm_quad(const int* qwptr, const unsigned short* qidx, const T* qval, const int* qrows, const T* diag,
                         const int* hptr, const int* hcol, int tile_rows, int hmax, const T* shift, const T* post, const int* xmap,
                         const int* ymap, const T* x, int64_t ldx, T* y, int64_t ldy, int64_t n, int ncols, const T* dot_with,
                         T* dot_out, void* dot_ws, cudaStream_t st) {
  constexpr int VEC = 16 / sizeof(T);
  constexpr int CW = 4 * VEC;
  MGP_CHECK_ARG(qwptr && qidx && qval && qrows && diag && hptr && hcol && x && y, "lap_spmm_quad: null pointer");
  MGP_CHECK_ARG(n > 0 && ncols > 0 && ldx >= ncols && ldy >= ncols && hmax >= 0, "lap_spmm_quad: bad shape");
  MGP_CHECK_ARG(x != y, "lap_spmm_quad: X and Y must not alias");
  MGP_CHECK_ARG((dot_out == nullptr) || (dot_with && dot_ws), "lap_spmm_quad: dot epilogue needs dot_with and dot_ws");
  const bool ok = tile_rows == kQdRows && (ncols % CW == 0) && (ldx % VEC == 0) && (ldy % VEC == 0) && (((uintptr_t)x) % 16 == 0) &&
                  (((uintptr_t)y) % 16 == 0) && (((uintptr_t)qval) % 16 == 0) && (dot_with == nullptr || ((uintptr_t)dot_with) % 16 == 0);
  if (!ok) return MGP_EUNSUPPORTED;
  const size_t smem = (size_t)(kQdRows + hmax + 4) * 64;
  if (smem > 200 * 1024) return MGP_EUNSUPPORTED;
  const int64_t ntiles = ceil_div(n, (int64_t)kQdRows);
  if (dot_out && (size_t)ntiles * CW * sizeof(T) > (size_t)kNumSMs * 8 * 32 * 8) return MGP_EUNSUPPORTED;   // partials must fit dot_ws
  auto kern = lap_spmm_quad_kernel<T>;
  static size_t configured = 0;
  if (smem > configured) {
    MGP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  unsigned int* counter = dot_out ? reinterpret_cast<unsigned int*>(dot_ws) : nullptr;
  T* partials = dot_out ? reinterpret_cast<T*>(reinterpret_cast<char*>(dot_ws) + 256) : nullptr;
  for (int c0 = 0; c0 < ncols; c0 += CW) {
    kern<<<(unsigned)ntiles, kQdThreads, smem, st>>>(qwptr, qidx, qval, qrows, diag, hptr, hcol, shift, post, xmap, ymap, x, ldx, y,
                                                    ldy, n, c0, dot_out ? dot_with : nullptr, dot_out, partials, counter,
                                                    (dot_out && dot_with == x) ? 1 : 0);
    MGP_LAUNCH_CHECK();
  }
  return MGP_OK;
}

}  // namespace mgp

extern "C" {

int mgp_lap_spmm_quad_f32(const int32_t* qwptr, const uint16_t* qidx, const float* qval, const int32_t* qrows, const float* diag,
                          const int32_t* hptr, const int32_t* hcol, int32_t tile_rows, int32_t hmax, const float* shift,
                          const float* post, const int32_t* xmap, const int32_t* ymap, const float* x, int64_t ldx, float* y,
                          int64_t ldy, int64_t n, int32_t ncols, const float* dot_with, float* dot_out, void* dot_ws, void* stream) {
  return mgp::lap_spmm_quad<float>(qwptr, qidx, qval, qrows, diag, hptr, hcol, tile_rows, hmax, shift, post, xmap, ymap, x, ldx, y,
                                   ldy, n, ncols, dot_with, dot_out, dot_ws, (cudaStream_t)stream);
}
int mgp_lap_spmm_quad_f64(const int32_t* qwptr, const uint16_t* qidx, const double* qval, const int32_t* qrows, const double* diag,
                          const int32_t* hptr, const int32_t* hcol, int32_t tile_rows, int32_t hmax, const double* shift,
                          const double* post, const int32_t* xmap, const int32_t* ymap, const double* x, int64_t ldx, double* y,
                          int64_t ldy, int64_t n, int32_t ncols, const double* dot_with, double* dot_out, void* dot_ws, void* stream) {
  return mgp::lap_spmm_quad<double>(qwptr, qidx, qval, qrows, diag, hptr, hcol, tile_rows, hmax, shift, post, xmap, ymap, x, ldx, y,
                                    ldy, n, ncols, dot_with, dot_out, dot_ws, (cudaStream_t)stream);
}

}  // extern "C"
