// Single-column SpMV on the warp-interleaved tile streams of lap_spmm_wi.cu:
//
//   y = post .* ( (diag + shift) .* x  -  A x )            x, y: one column (stride ldx / ldy between rows)
//
// Replaces graph_laplacian_operator.py:108-124 (2 x torch_sparse.spmm + diagonal) / one step of
// precision_matern_operator.py:28-32 for a single right-hand side -- what Lanczos and single-RHS CG call.
//
// Why a separate kernel: with one column an X "row" is 4 bytes, so the 64-byte-row machinery of the C = 16 kernel (TMA
// stages, producer warps) has nothing to hide -- the kernel is a pure stream of (16-bit index, value) pairs, 6 bytes per
// nonzero against the 8 of a CSR walk, and the tile's slice of x (own 128 rows + ~240 halo rows, < 3 KB) sits in shared
// memory so that no gather ever goes through L1.  One 256-thread block per tile, ~8 blocks per SM: the memory parallelism
// comes from occupancy (each thread keeps 4 (index, value) loads in flight), not from an explicit pipeline.
// Stream layout (graph.py): for warp block w of tile t, step s, lane l: position wptr[16 t + w] + 32 s + l holds nonzero
// 4 s + (l & 3) of row 8 w + (l >> 2); padding entries have value 0 and a valid index.
#include "common.cuh"
#include "pipe_common.cuh"
#include "spmm_common.cuh"

namespace mgp {

constexpr int kSvRows = 128;
constexpr int kSvThreads = 256;
// capacity of the partial-sum area of the dot workspace (mgp_lap_spmm_dot_ws_bytes = 256 + kNumSMs * 8 * 32 * 8 bytes)
constexpr size_t kDotWsPartialBytes = (size_t)kNumSMs * 8 * 32 * 8;

template <typename T>
__global__ void __launch_bounds__(kSvThreads)
lap_spmv_tile_kernel(const int* __restrict__ wptr, const unsigned short* __restrict__ wcol, const T* __restrict__ aw,
                     const T* __restrict__ diag, const int* __restrict__ hptr, const int* __restrict__ hcol,
                     const T* __restrict__ shift_p, const T* __restrict__ post, const int* __restrict__ xmap,
                     const int* __restrict__ ymap, const T* __restrict__ x, int64_t ldx, T* __restrict__ y, int64_t ldy,
                     int64_t n, const T* __restrict__ dot_with, T* __restrict__ dot_out, T* __restrict__ partials,
                     unsigned int* __restrict__ counter, int dot_is_x) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* xs = reinterpret_cast<T*>(smem_raw);                 // [128 own rows | halo rows]
  T dsum = T(0);                                          // this thread's share of dot_with^T y (optional epilogue)
  const int t = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t row0 = (int64_t)t * kSvRows;
  const int nrows = (int)min((int64_t)kSvRows, n - row0);
  const int h0 = hptr[t], nh = hptr[t + 1] - h0;
  for (int r = tid; r < nrows; r += kSvThreads) {
    const int64_t src = xmap ? (int64_t)__ldg(xmap + row0 + r) : row0 + r;
    xs[r] = __ldg(x + src * ldx);
  }
  for (int h = tid; h < nh; h += kSvThreads) {
    int64_t src = __ldg(hcol + h0 + h);
    if (xmap) src = __ldg(xmap + src);
    xs[kSvRows + h] = __ldg(x + src * ldx);
  }
  __syncthreads();
  const T shift = shift_p ? *shift_p : T(0);
  for (int wb = warp; wb < 16; wb += kSvThreads / 32) {
    const int base = wptr[16 * t + wb];
    const int steps = (wptr[16 * t + wb + 1] - base) >> 5;
    const unsigned short* cp = wcol + base + lane;
    const T* vp = aw + base + lane;
    T acc0 = T(0), acc1 = T(0);
    int s = 0;
    for (; s + 4 <= steps; s += 4) {                       // 4 (index, value) pairs in flight per thread
      const unsigned short j0 = __ldcs(cp), j1 = __ldcs(cp + 32), j2 = __ldcs(cp + 64), j3 = __ldcs(cp + 96);
      const T a0 = __ldcs(vp), a1 = __ldcs(vp + 32), a2 = __ldcs(vp + 64), a3 = __ldcs(vp + 96);
      acc0 = fma(a0, xs[j0], acc0); acc1 = fma(a1, xs[j1], acc1);
      acc0 = fma(a2, xs[j2], acc0); acc1 = fma(a3, xs[j3], acc1);
      cp += 128; vp += 128;
    }
    for (; s < steps; ++s) {
      acc0 = fma(__ldcs(vp), xs[__ldcs(cp)], acc0);
      cp += 32; vp += 32;
    }
    T acc = acc0 + acc1;
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    const int r = wb * 8 + (lane >> 2);
    if ((lane & 3) == 0 && r < nrows) {
      const int64_t row = row0 + r;
      const T d = __ldg(diag + row) + shift;
      T out = d * xs[r] - acc;
      if (post) out *= __ldg(post + row);
      const int64_t yrow = ymap ? (int64_t)__ldg(ymap + row) : row;
      y[yrow * ldy] = out;
      if (dot_out) {
        T dw = xs[r];
        if (!dot_is_x) {
          const int64_t drow = xmap ? (int64_t)__ldg(xmap + row) : row;
          dw = __ldg(dot_with + drow * ldx);
        }
        dsum = fma(dw, out, dsum);
      }
    }
  }
  if (dot_out) {
    // block sum -> partials[block]; the last block to arrive adds the partials in a fixed order (deterministic)
    __shared__ T red[kSvThreads];
    dsum = warp_sum(dsum);
    __syncthreads();
    if (lane == 0) red[warp] = dsum;
    __syncthreads();
    if (tid == 0) {
      T s = T(0);
#pragma unroll
      for (int w = 0; w < kSvThreads / 32; ++w) s += red[w];
      partials[blockIdx.x] = s;
    }
    if (last_block_ticket(counter)) {
      T s = T(0);
      for (int b = tid; b < (int)gridDim.x; b += kSvThreads) s += __ldcg(partials + b);
      red[tid] = s;
      __syncthreads();
      if (tid == 0) {
        T tot = T(0);
        for (int i = 0; i < kSvThreads; ++i) tot += red[i];
        dot_out[0] = tot;
      }
    }
  }
}

template <typename T>
static int lap_spmv_tile(const int* wptr, const unsigned short* wcol, const T* aw, const T* diag, const int* hptr, const int* hcol,
                         int tile_rows, int hmax, const T* shift, const T* post, const int* xmap, const int* ymap, const T* x,
                         int64_t ldx, T* y, int64_t ldy, int64_t n, const T* dot_with, T* dot_out, void* dot_ws, cudaStream_t st) {
  MGP_CHECK_ARG(wptr && wcol && aw && diag && hptr && hcol && x && y, "lap_spmv_tile: null pointer");
  MGP_CHECK_ARG((dot_out == nullptr) || (dot_with && dot_ws), "lap_spmv_tile: dot epilogue needs dot_with and dot_ws");
  MGP_CHECK_ARG(n > 0 && ldx >= 1 && ldy >= 1 && hmax >= 0, "lap_spmv_tile: bad shape");
  MGP_CHECK_ARG(x != y, "lap_spmv_tile: X and Y must not alias");
  if (tile_rows != kSvRows) return MGP_EUNSUPPORTED;
  const size_t smem = (size_t)(kSvRows + hmax + 4) * sizeof(T);
  if (smem > 48 * 1024) return MGP_EUNSUPPORTED;
  const int64_t ntiles = ceil_div(n, (int64_t)kSvRows);
  if (dot_out && (size_t)ntiles * sizeof(T) > kDotWsPartialBytes) return MGP_EUNSUPPORTED;   // one partial per tile must fit dot_ws
  unsigned int* counter = dot_out ? reinterpret_cast<unsigned int*>(dot_ws) : nullptr;
  T* partials = dot_out ? reinterpret_cast<T*>(reinterpret_cast<char*>(dot_ws) + 256) : nullptr;
  lap_spmv_tile_kernel<T><<<(unsigned)ntiles, kSvThreads, smem, st>>>(wptr, wcol, aw, diag, hptr, hcol, shift, post, xmap, ymap, x,
                                                                       ldx, y, ldy, n, dot_out ? dot_with : nullptr, dot_out,
                                                                       partials, counter, (dot_out && dot_with == x) ? 1 : 0);
  MGP_LAUNCH_CHECK();
  return MGP_OK;
}

}  // namespace mgp

extern "C" {

int mgp_lap_spmv_tile_f32(const int32_t* wptr, const uint16_t* wcol, const float* aw, const float* diag, const int32_t* hptr,
                          const int32_t* hcol, int32_t tile_rows, int32_t hmax, const float* shift, const float* post,
                          const int32_t* xmap, const int32_t* ymap, const float* x, int64_t ldx, float* y, int64_t ldy, int64_t n,
                          const float* dot_with, float* dot_out, void* dot_ws, void* stream) {
  return mgp::lap_spmv_tile<float>(wptr, wcol, aw, diag, hptr, hcol, tile_rows, hmax, shift, post, xmap, ymap, x, ldx, y, ldy, n,
                                   dot_with, dot_out, dot_ws, (cudaStream_t)stream);
}
int mgp_lap_spmv_tile_f64(const int32_t* wptr, const uint16_t* wcol, const double* aw, const double* diag, const int32_t* hptr,
                          const int32_t* hcol, int32_t tile_rows, int32_t hmax, const double* shift, const double* post,
                          const int32_t* xmap, const int32_t* ymap, const double* x, int64_t ldx, double* y, int64_t ldy, int64_t n,
                          const double* dot_with, double* dot_out, void* dot_ws, void* stream) {
  return mgp::lap_spmv_tile<double>(wptr, wcol, aw, diag, hptr, hcol, tile_rows, hmax, shift, post, xmap, ymap, x, ldx, y, ldy, n,
                                   dot_with, dot_out, dot_ws, (cudaStream_t)stream);
}

}  // extern "C"
