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

// ---- the same shape for 64-byte rows (16 fp32 / 8 fp64 columns per pass): one block per tile, no producer warps ------------
// The tile's X rows are staged with 16-byte cp.async by all 256 threads, the (index, value) streams are read straight from
// global memory (coalesced, 4 steps in flight per warp); overlap comes from the ~5 blocks resident per SM.  Row walk and
// lane <-> chunk rotation are those of lap_spmm_wi_kernel (same stream layout, same bank behaviour).
template <typename T>
__global__ void __launch_bounds__(kSvThreads)
lap_spmm_tile64_kernel(const int* __restrict__ wptr, const unsigned short* __restrict__ wcol, const T* __restrict__ aw,
                       const T* __restrict__ diag, const int* __restrict__ hptr, const int* __restrict__ hcol,
                       const T* __restrict__ shift_p, const T* __restrict__ post, const int* __restrict__ xmap,
                       const int* __restrict__ ymap, const T* __restrict__ x, int64_t ldx, T* __restrict__ y, int64_t ldy,
                       int64_t n, int c0, const T* __restrict__ dot_with, T* __restrict__ dot_out, T* __restrict__ partials,
                       unsigned int* __restrict__ counter, int dot_is_x) {
  constexpr int VEC = 16 / sizeof(T);
  constexpr int CW = 4 * VEC;
  constexpr int ROW_BYTES = 64;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char* const xs = smem_raw;
  const int t = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t row0 = (int64_t)t * kSvRows;
  const int nrows = (int)min((int64_t)kSvRows, n - row0);
  const int h0 = hptr[t], nh = hptr[t + 1] - h0;
  {
    const int ch = tid & 3;
    const unsigned char* xb = reinterpret_cast<const unsigned char*>(x + c0) + ch * 16;
    const int64_t ldxb = ldx * (int64_t)sizeof(T);
    for (int rr = tid >> 2; rr < nrows + nh; rr += kSvThreads / 4) {
      int64_t src = rr < nrows ? row0 + rr : (int64_t)__ldg(hcol + h0 + rr - nrows);
      if (xmap) src = __ldg(xmap + src);
      const int dst = rr < nrows ? rr : kSvRows + (rr - nrows);
      cp_async16(xs + (size_t)dst * ROW_BYTES + ch * 16, xb + src * ldxb);
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
  }
  __syncthreads();
  const T shift = shift_p ? *shift_p : T(0);
  const int slot = lane >> 2, l = lane & 3;
  const int cbase = c0 + l * VEC;
  const uint32_t o0 = (uint32_t)(((0 + l) & 3) * 16), o1 = (uint32_t)(((1 + l) & 3) * 16);
  const uint32_t o2 = (uint32_t)(((2 + l) & 3) * 16), o3 = (uint32_t)(((3 + l) & 3) * 16);
  T dsum[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) dsum[v] = T(0);
  for (int wb = warp; wb < 16; wb += kSvThreads / 32) {
    const int base = wptr[16 * t + wb];
    const int steps = (wptr[16 * t + wb + 1] - base) >> 5;
    const unsigned short* cp = wcol + base + lane;
    const T* vp = aw + base + lane;
    T acc[4][VEC];
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
      for (int v = 0; v < VEC; ++v) acc[c][v] = T(0);
    auto fma_row = [&](uint32_t j, T wv) {
      const unsigned char* xr = xs + j * ROW_BYTES;
      const Vec<T, VEC> x0 = *reinterpret_cast<const Vec<T, VEC>*>(xr + o0);
      const Vec<T, VEC> x1 = *reinterpret_cast<const Vec<T, VEC>*>(xr + o1);
      const Vec<T, VEC> x2 = *reinterpret_cast<const Vec<T, VEC>*>(xr + o2);
      const Vec<T, VEC> x3 = *reinterpret_cast<const Vec<T, VEC>*>(xr + o3);
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        acc[0][v] = fma(wv, x0.v[v], acc[0][v]);
        acc[1][v] = fma(wv, x1.v[v], acc[1][v]);
        acc[2][v] = fma(wv, x2.v[v], acc[2][v]);
        acc[3][v] = fma(wv, x3.v[v], acc[3][v]);
      }
    };
    int s = 0;
    for (; s + 4 <= steps; s += 4) {
      const uint32_t j0 = __ldcs(cp), j1 = __ldcs(cp + 32), j2 = __ldcs(cp + 64), j3 = __ldcs(cp + 96);
      const T a0 = __ldcs(vp), a1 = __ldcs(vp + 32), a2 = __ldcs(vp + 64), a3 = __ldcs(vp + 96);
      fma_row(j0, a0); fma_row(j1, a1); fma_row(j2, a2); fma_row(j3, a3);
      cp += 128; vp += 128;
    }
    for (; s < steps; ++s) {
      fma_row(__ldcs(cp), __ldcs(vp));
      cp += 32; vp += 32;
    }
    // lane l ends up with chunk l complete: its own acc[0] plus acc[4 - d] of the lane d places further (mod 4) in the slot
    T res[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) res[v] = acc[0][v];
#pragma unroll
    for (int d = 1; d < 4; ++d) {
      const int src = (lane & ~3) | ((l + d) & 3);
#pragma unroll
      for (int v = 0; v < VEC; ++v) res[v] += __shfl_sync(0xffffffffu, acc[4 - d][v], src);
    }
    const int r = wb * 8 + slot;
    if (r < nrows) {
      const int64_t row = row0 + r;
      const Vec<T, VEC> xi = *reinterpret_cast<const Vec<T, VEC>*>(xs + (size_t)r * ROW_BYTES + l * 16);
      const T d = __ldg(diag + row) + shift;
      const T po = post ? __ldg(post + row) : T(1);
      Vec<T, VEC> out;
#pragma unroll
      for (int v = 0; v < VEC; ++v) out.v[v] = po * (d * xi.v[v] - res[v]);
      const int64_t yrow = ymap ? (int64_t)__ldg(ymap + row) : row;
      st_vec<T, VEC>(y + yrow * ldy + cbase, out);
      if (dot_out) {
        Vec<T, VEC> dw = xi;
        if (!dot_is_x) {
          const int64_t drow = xmap ? (int64_t)__ldg(xmap + row) : row;
          dw = ldg_vec<T, VEC>(dot_with + drow * ldx + cbase);
        }
#pragma unroll
        for (int v = 0; v < VEC; ++v) dsum[v] = fma(dw.v[v], out.v[v], dsum[v]);
      }
    }
  }
  if (dot_out) {
    __syncthreads();
    spmm_dot_epilogue<T, VEC, 4, CW, kSvThreads>(dsum, CW, c0, partials, counter, dot_out);
  }
}

template <typename T>
static int lap_spmm_tile64(const int* wptr, const unsigned short* wcol, const T* aw, const T* diag, const int* hptr,
                           const int* hcol, int tile_rows, int hmax, const T* shift, const T* post, const int* xmap,
                           const int* ymap, const T* x, int64_t ldx, T* y, int64_t ldy, int64_t n, int ncols, const T* dot_with,
                           T* dot_out, void* dot_ws, cudaStream_t st) {
  constexpr int VEC = 16 / sizeof(T);
  constexpr int CW = 4 * VEC;
  MGP_CHECK_ARG(wptr && wcol && aw && diag && hptr && hcol && x && y, "lap_spmm_tile64: null pointer");
  MGP_CHECK_ARG(n > 0 && ncols > 0 && ldx >= ncols && ldy >= ncols && hmax >= 0, "lap_spmm_tile64: bad shape");
  MGP_CHECK_ARG(x != y, "lap_spmm_tile64: X and Y must not alias");
  MGP_CHECK_ARG((dot_out == nullptr) || (dot_with && dot_ws), "lap_spmm_tile64: dot epilogue needs dot_with and dot_ws");
  const bool ok = tile_rows == kSvRows && (ncols % CW == 0) && (ldx % VEC == 0) && (ldy % VEC == 0) && (((uintptr_t)x) % 16 == 0) &&
                  (((uintptr_t)y) % 16 == 0) && (dot_with == nullptr || ((uintptr_t)dot_with) % 16 == 0);
  if (!ok) return MGP_EUNSUPPORTED;
  const size_t smem = (size_t)(kSvRows + hmax + 4) * 64;
  if (smem > 200 * 1024) return MGP_EUNSUPPORTED;
  auto kern = lap_spmm_tile64_kernel<T>;
  static size_t configured = 0;
  if (smem > configured) {
    MGP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  const int64_t ntiles = ceil_div(n, (int64_t)kSvRows);
  if (dot_out && (size_t)ntiles * CW * sizeof(T) > kDotWsPartialBytes) return MGP_EUNSUPPORTED;   // per-tile partials must fit dot_ws
  unsigned int* counter = dot_out ? reinterpret_cast<unsigned int*>(dot_ws) : nullptr;
  T* partials = dot_out ? reinterpret_cast<T*>(reinterpret_cast<char*>(dot_ws) + 256) : nullptr;
  for (int c0 = 0; c0 < ncols; c0 += CW) {
    kern<<<(unsigned)ntiles, kSvThreads, smem, st>>>(wptr, wcol, aw, diag, hptr, hcol, shift, post, xmap, ymap, x, ldx, y, ldy, n, c0,
                                                    dot_out ? dot_with : nullptr, dot_out, partials, counter,
                                                    (dot_out && dot_with == x) ? 1 : 0);
    MGP_LAUNCH_CHECK();
  }
  return MGP_OK;
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

int mgp_lap_spmm_tile64_f32(const int32_t* wptr, const uint16_t* wcol, const float* aw, const float* diag, const int32_t* hptr,
                            const int32_t* hcol, int32_t tile_rows, int32_t hmax, const float* shift, const float* post,
                            const int32_t* xmap, const int32_t* ymap, const float* x, int64_t ldx, float* y, int64_t ldy, int64_t n,
                            int32_t ncols, const float* dot_with, float* dot_out, void* dot_ws, void* stream) {
  return mgp::lap_spmm_tile64<float>(wptr, wcol, aw, diag, hptr, hcol, tile_rows, hmax, shift, post, xmap, ymap, x, ldx, y, ldy, n,
                                     ncols, dot_with, dot_out, dot_ws, (cudaStream_t)stream);
}
int mgp_lap_spmm_tile64_f64(const int32_t* wptr, const uint16_t* wcol, const double* aw, const double* diag, const int32_t* hptr,
                            const int32_t* hcol, int32_t tile_rows, int32_t hmax, const double* shift, const double* post,
                            const int32_t* xmap, const int32_t* ymap, const double* x, int64_t ldx, double* y, int64_t ldy, int64_t n,
                            int32_t ncols, const double* dot_with, double* dot_out, void* dot_ws, void* stream) {
  return mgp::lap_spmm_tile64<double>(wptr, wcol, aw, diag, hptr, hcol, tile_rows, hmax, shift, post, xmap, ymap, x, ldx, y, ldy, n,
                                      ncols, dot_with, dot_out, dot_ws, (cudaStream_t)stream);
}

}  // extern "C"
