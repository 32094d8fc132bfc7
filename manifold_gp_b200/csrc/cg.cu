// Batched conjugate gradients: the fused vector kernels of mBCG.
//
// Algorithm: linear_operator.utils.linear_cg (third party; restated in oracle/solvers.py; reference call sites
// utils/train_model.py:55,67-68, operators/precision_matern_operator.py:53, operators/schur_complement_operator.py:28).
// linear_operator runs ~10 un-fused elementwise / reduction launches per iteration and reads the residual norm
// back to the host every iteration.  Here an iteration is
//     [operator matvec, optionally with the p^T A p epilogue]  ->  cg_update  ->  cg_pupdate
// all scalars (alpha, beta, norms, masks, iteration counter, convergence flag) live in a device `state` block,
// reductions are deterministic (per-block partials + last-block fixed-order reduce), and every kernel is a no-op
// once the done flag is set, so the host enqueues iterations in chunks and polls one flag.
//
// State layout (elements of T, C = ncols):  8 arrays of C then 16 scalars -- see CgState below / mgp_b200.h.
#include "common.cuh"

namespace mgp {

enum : int { S_RHSNORM = 0, S_RZ, S_PAP, S_ALPHA, S_BETA, S_RESID, S_RHSZERO, S_CONV, S_NARR };
enum : int { K_MEAN = 0, K_DONE, K_ITER, K_TOL, K_EPS, K_STOP, K_MINITER, K_NTRIMIN, K_MAXITER, K_XPEND, K_NSCAL = 16 };

constexpr int kCgBlock = 256;
constexpr int kCgMaxCols = 128;
constexpr int kCgMaxCB = kCgMaxCols / 32;

static inline int cg_grid(int64_t n, int rows_per_block) {
  int64_t b = ceil_div(n, rows_per_block);
  const int64_t cap = (int64_t)kNumSMs * 8;
  if (b > cap) b = cap;
  return (int)(b < 1 ? 1 : b);
}

static inline int col_lanes(int ncols) {
  int cp = 1;
  while (cp < ncols && cp < 32) cp <<= 1;
  return cp;
}

// Block-level per-column reduction of `acc[cb]` (column = cb*CP + tid%CP) into partials[blockIdx.x*ncols + c].
template <typename T>
__device__ __forceinline__ void block_col_reduce(const T* acc, int ncb, int CP, int ncols, T* partials) {
  __shared__ T sm[kCgBlock / 32][32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cl = tid % CP;  // CP divides 32, so cl == lane % CP
  for (int cb = 0; cb < ncb; ++cb) {
    T v = acc[cb];
    v = subwarp_sum(v, CP, 32);
    __syncthreads();
    if (lane < CP) sm[warp][lane] = v;
    __syncthreads();
    if (tid < CP) {
      T s = T(0);
#pragma unroll
      for (int w = 0; w < kCgBlock / 32; ++w) s += sm[w][tid];
      const int c = cb * CP + cl;
      if (c < ncols) partials[(int64_t)blockIdx.x * ncols + c] = s;
    }
  }
}

// Fixed-order reduction of partials over blocks by the calling (last) block; result for column c in smem out[c].
// The loads of a thread are independent (8 in flight): the r01 version chained ~74 dependent L2 round trips per thread
// (~10 us of serial tail in every r-update launch, ncu launch list profiles/r01_*).
template <typename T>
__device__ __forceinline__ void last_block_reduce(const T* partials, int ncols, T* out /* smem [kCgMaxCols] */) {
  __shared__ T red[kCgBlock];
  const int tid = threadIdx.x;
  const int nb = (int)gridDim.x;
  for (int cbase = 0; cbase < ncols; cbase += 32) {
    const int cw = min(32, ncols - cbase);
    const int per = kCgBlock / cw;
    const int c = tid % cw, r = tid / cw;
    T s = T(0);
    if (r < per) {
      const T* q = partials + cbase + c;
      int b = r;
      for (; b + 7 * per < nb; b += 8 * per) {
        T t[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) t[u] = __ldcg(q + (int64_t)(b + u * per) * ncols);
#pragma unroll
        for (int u = 0; u < 8; ++u) s += t[u];
      }
      for (; b < nb; b += per) s += __ldcg(q + (int64_t)b * ncols);
    }
    __syncthreads();
    red[tid] = (r < per) ? s : T(0);
    __syncthreads();
    if (tid < cw) {
      T t = T(0);
      for (int rr = 0; rr < per; ++rr) t += red[rr * cw + tid];
      out[cbase + tid] = t;
    }
  }
  __syncthreads();
}

// sum of one value per thread over the block, result valid in thread 0 (fixed order: lanes by shuffle tree, then warps 0..7)
template <typename T>
__device__ __forceinline__ T block_sum_t0(T v) {
  __shared__ T wsum[kCgBlock / 32];
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = v;
  __syncthreads();
  T s = T(0);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w = 0; w < kCgBlock / 32; ++w) s += wsum[w];
  }
  return s;
}

// per-thread accumulators of the vectorised kernels (thread t holds columns (t*N) % ld .. +N-1) -> one partial per column of
// the block: lanes with equal columns are combined by a shuffle tree, the 8 warps through shared memory (fixed order)
template <typename T, int N>
__device__ __forceinline__ void vec_block_partials(const T (&acc)[N], int ld, int ncols, T* dst /* partials of this block */,
                                                   int col0 = -1 /* first column held by this thread; default (tid * N) % ld */) {
  __shared__ T wpart[kCgBlock / 32][kCgMaxCols];
  __shared__ int wblock[kCgBlock / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int groups = ld / N;                       // threads per row of the vector; a power of two
  const int span = 32 * N;                         // columns a warp covers when groups >= 32 (one aligned block of the row)
  if (col0 < 0) col0 = (tid * N) % ld;             // threads whose indices are congruent mod `groups` hold the same columns
  T a[N];
#pragma unroll
  for (int u = 0; u < N; ++u) {
    a[u] = acc[u];
    for (int o = groups; o < 32; o <<= 1) a[u] += __shfl_xor_sync(0xffffffffu, a[u], o);
  }
  if (groups >= 32) {
    if (lane == 0) wblock[warp] = col0 / span;
#pragma unroll
    for (int u = 0; u < N; ++u) wpart[warp][col0 + u] = a[u];
  } else if (lane < groups) {
#pragma unroll
    for (int u = 0; u < N; ++u) wpart[warp][col0 + u] = a[u];
  }
  __syncthreads();
  if (tid < ncols) {
    T s = T(0);
    if (groups >= 32) {
      for (int w = 0; w < kCgBlock / 32; ++w)
        if (wblock[w] == tid / span) s += wpart[w][tid];
    } else {
#pragma unroll
      for (int w = 0; w < kCgBlock / 32; ++w) s += wpart[w][tid];
    }
    dst[tid] = s;
  }
}

template <typename T>
struct CgWs {
  unsigned int* counter;
  T* partials;
};
template <typename T>
__host__ __device__ inline CgWs<T> cg_ws(void* ws) {
  CgWs<T> w;
  w.counter = reinterpret_cast<unsigned int*>(ws);
  w.partials = reinterpret_cast<T*>(reinterpret_cast<char*>(ws) + 256);
  return w;
}

// ---- scalar bookkeeping shared by the single-GPU kernels (run by the last block) and the multi-GPU split kernels (run
//      by a one-block kernel after the all-reduce of the column sums).  `tot` is in shared memory, whole block calls. ----
template <typename T>
__device__ __forceinline__ void cg_finish_norm(T* state, const T* tot, int ncols, T eps) {
  for (int c = threadIdx.x; c < ncols; c += kCgBlock) {
    T nrm = dev_sqrt<T>(tot[c]);
    const bool zero = nrm < eps;
    state[S_RHSZERO * ncols + c] = zero ? T(1) : T(0);
    state[S_RHSNORM * ncols + c] = zero ? T(1) : nrm;
  }
}

template <typename T>
__device__ __forceinline__ void cg_finish_init(T* state, const T* tot, int ncols, T tol, T eps, T stop, int max_iter,
                                               int n_tridiag_iter) {
  const int tid = threadIdx.x;
  T local = T(0);
  int all_conv = 1;
  for (int c = tid; c < ncols; c += kCgBlock) {
    const T rn = dev_sqrt<T>(tot[c]);
    state[S_RZ * ncols + c] = tot[c];
    state[S_RESID * ncols + c] = rn;
    state[S_CONV * ncols + c] = (rn < stop) ? T(1) : T(0);
    state[S_PAP * ncols + c] = T(0);
    state[S_ALPHA * ncols + c] = T(0);
    state[S_BETA * ncols + c] = T(0);
    local += rn;
    if (!(rn < stop)) all_conv = 0;
  }
  const int any_unconv = __syncthreads_or(!all_conv);
  const T s = block_sum_t0<T>(local);
  if (tid == 0) {
    T* k = state + S_NARR * ncols;
    k[K_MEAN] = s / T(ncols);
    // published: "if has_converged.all() and not n_tridiag: n_iter = 0"
    k[K_DONE] = (!any_unconv && n_tridiag_iter == 0) ? T(1) : T(0);
    k[K_ITER] = T(0);
    k[K_TOL] = tol;
    k[K_EPS] = eps;
    k[K_STOP] = stop;
    k[K_MINITER] = T(min(10, max_iter - 1));
    k[K_NTRIMIN] = T(n_tridiag_iter > 0 ? min(n_tridiag_iter, max_iter - 1) : 0);
    k[K_MAXITER] = T(max_iter);
    if (max_iter <= 0) k[K_DONE] = T(2);
  }
}

template <typename T>
__device__ __forceinline__ T cg_alpha_of(const T* state, int ncols, int c, T eps);

template <typename T>
__device__ __forceinline__ void cg_finish_update(T* state, const T* tot, int ncols, T* hist, int max_hist) {
  const int tid = threadIdx.x;
  T* k = state + S_NARR * ncols;
  const T eps = k[K_EPS];
  const int it = (int)k[K_ITER];
  const T stop = k[K_STOP];
  T local = T(0);
  for (int c = tid; c < ncols; c += kCgBlock) {
    const T aa = cg_alpha_of<T>(state, ncols, c, eps);
    const T rz_old = state[S_RZ * ncols + c];
    const T rz_new = tot[c];
    const bool bz = rz_old < eps;
    const T beta = bz ? T(0) : rz_new / rz_old;
    T rn = dev_sqrt<T>(rz_new);
    if (state[S_RHSZERO * ncols + c] != T(0)) rn = T(0);
    state[S_ALPHA * ncols + c] = aa;
    state[S_BETA * ncols + c] = beta;
    state[S_RZ * ncols + c] = rz_new;
    state[S_RESID * ncols + c] = rn;
    state[S_CONV * ncols + c] = (rn < stop) ? T(1) : T(0);
    if (hist && it < max_hist) {
      hist[((int64_t)it * 2 + 0) * ncols + c] = aa;
      hist[((int64_t)it * 2 + 1) * ncols + c] = beta;
    }
    local += rn;
  }
  const T s = block_sum_t0<T>(local);
  if (tid == 0) {
    const T mean = s / T(ncols);
    k[K_MEAN] = mean;
    k[K_ITER] = T(it + 1);
    k[K_XPEND] = T(1);      // x += alpha p of this iteration is still to be applied (split-update path: cg_pxupdate)
    // published stopping rule: k >= min(10, max_iter-1) and mean residual < tol and the tridiagonal has enough rows
    const bool tri_pending = (k[K_NTRIMIN] > T(0)) && (T(it) < k[K_NTRIMIN]);
    if (T(it) >= k[K_MINITER] && mean < k[K_TOL] && !tri_pending) k[K_DONE] = T(1);
    else if (T(it + 1) >= k[K_MAXITER]) k[K_DONE] = T(2);
  }
}

// tot (shared) -> rbuf (global): the multi-GPU path all-reduces rbuf before the scalars are finished
template <typename T>
__device__ __forceinline__ void cg_export_tot(const T* tot, T* rbuf, int ncols) {
  for (int c = threadIdx.x; c < ncols; c += kCgBlock) rbuf[c] = tot[c];
}

// ---- init pass 1: column norms of b ---------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kCgBlock)
cg_norm_kernel(const T* __restrict__ b, int64_t ldb, int64_t n, int ncols, int CP, T eps, T* __restrict__ state,
               void* ws, T* __restrict__ rbuf) {
  CgWs<T> w = cg_ws<T>(ws);
  const int tid = threadIdx.x;
  const int cl = tid % CP, rl = tid / CP, rpb = kCgBlock / CP;
  const int ncb = (ncols + CP - 1) / CP;
  T acc[kCgMaxCB];
#pragma unroll
  for (int i = 0; i < kCgMaxCB; ++i) acc[i] = T(0);
  for (int64_t row = (int64_t)blockIdx.x * rpb + rl; row < n; row += (int64_t)gridDim.x * rpb) {
#pragma unroll
    for (int cb = 0; cb < kCgMaxCB; ++cb) {
      const int c = cb * CP + cl;
      if (cb < ncb && c < ncols) {
        const T v = b[row * ldb + c];
        acc[cb] = fma(v, v, acc[cb]);
      }
    }
  }
  block_col_reduce<T>(acc, ncb, CP, ncols, w.partials);
  if (last_block_ticket(w.counter)) {
    __shared__ T tot[kCgMaxCols];
    last_block_reduce<T>(w.partials, ncols, tot);
    if (rbuf) cg_export_tot<T>(tot, rbuf, ncols);
    else cg_finish_norm<T>(state, tot, ncols, eps);
  }
}

// ---- init pass 2: r = b/|b|, p = r, x = 0, rz = |r|^2 ----------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kCgBlock)
cg_init_kernel(const T* __restrict__ b, int64_t ldb, T* __restrict__ x, T* __restrict__ r, T* __restrict__ p,
               int64_t ld, int64_t n, int ncols, int CP, T tol, T eps, T stop, int max_iter, int n_tridiag_iter,
               T* __restrict__ state, void* ws, T* __restrict__ rbuf) {
  CgWs<T> w = cg_ws<T>(ws);
  const int tid = threadIdx.x;
  const int cl = tid % CP, rl = tid / CP, rpb = kCgBlock / CP;
  const int ncb = (ncols + CP - 1) / CP;
  T acc[kCgMaxCB], inv[kCgMaxCB];
#pragma unroll
  for (int cb = 0; cb < kCgMaxCB; ++cb) {
    acc[cb] = T(0);
    const int c = cb * CP + cl;
    inv[cb] = (cb < ncb && c < ncols) ? state[S_RHSNORM * ncols + c] : T(1);
  }
  for (int64_t row = (int64_t)blockIdx.x * rpb + rl; row < n; row += (int64_t)gridDim.x * rpb) {
#pragma unroll
    for (int cb = 0; cb < kCgMaxCB; ++cb) {
      const int c = cb * CP + cl;
      if (cb < ncb && c < ncols) {
        const T v = b[row * ldb + c] / inv[cb];
        r[row * ld + c] = v;
        p[row * ld + c] = v;
        x[row * ld + c] = T(0);
        acc[cb] = fma(v, v, acc[cb]);
      }
    }
    // zero the padding columns so vectorised consumers never see garbage
    for (int c = ncols + tid % CP; c < ld; c += CP) {
      r[row * ld + c] = T(0); p[row * ld + c] = T(0); x[row * ld + c] = T(0);
    }
  }
  block_col_reduce<T>(acc, ncb, CP, ncols, w.partials);
  if (last_block_ticket(w.counter)) {
    __shared__ T tot[kCgMaxCols];
    last_block_reduce<T>(w.partials, ncols, tot);
    if (rbuf) cg_export_tot<T>(tot, rbuf, ncols);
    else cg_finish_init<T>(state, tot, ncols, tol, eps, stop, max_iter, n_tridiag_iter);
  }
}

// ---- pAp reduction (only when the matvec did not fuse it) ------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kCgBlock)
cg_pap_kernel(const T* __restrict__ p, const T* __restrict__ v, int64_t ld, int64_t n, int ncols, int CP,
              T* __restrict__ state, void* ws) {
  if (state[S_NARR * ncols + K_DONE] != T(0)) return;
  CgWs<T> w = cg_ws<T>(ws);
  const int tid = threadIdx.x;
  const int cl = tid % CP, rl = tid / CP, rpb = kCgBlock / CP;
  const int ncb = (ncols + CP - 1) / CP;
  T acc[kCgMaxCB];
#pragma unroll
  for (int i = 0; i < kCgMaxCB; ++i) acc[i] = T(0);
  for (int64_t row = (int64_t)blockIdx.x * rpb + rl; row < n; row += (int64_t)gridDim.x * rpb) {
#pragma unroll
    for (int cb = 0; cb < kCgMaxCB; ++cb) {
      const int c = cb * CP + cl;
      if (cb < ncb && c < ncols) acc[cb] = fma(p[row * ld + c], v[row * ld + c], acc[cb]);
    }
  }
  block_col_reduce<T>(acc, ncb, CP, ncols, w.partials);
  if (last_block_ticket(w.counter)) {
    __shared__ T tot[kCgMaxCols];
    last_block_reduce<T>(w.partials, ncols, tot);
    for (int c = tid; c < ncols; c += kCgBlock) state[S_PAP * ncols + c] = tot[c];
  }
}

// alpha_k from (rz, pAp, masks): the published "safe division"
template <typename T>
__device__ __forceinline__ T cg_alpha_of(const T* state, int ncols, int c, T eps) {
  const T pap = state[S_PAP * ncols + c];
  const T rz = state[S_RZ * ncols + c];
  const bool is_zero = pap < eps;
  T alpha = is_zero ? T(0) : rz / pap;
  if (state[S_CONV * ncols + c] != T(0)) alpha = T(0);
  return alpha;
}

// ---- x += alpha p ; r -= alpha v ; rz' = |r|^2 ; beta, norms, flags, history ------------------------------------
template <typename T>
__global__ void __launch_bounds__(kCgBlock)
cg_update_kernel(T* __restrict__ x, T* __restrict__ r, const T* __restrict__ p, const T* __restrict__ v, int64_t ld,
                 int64_t n, int ncols, int CP, T* __restrict__ state, T* __restrict__ hist, int max_hist, void* ws,
                 T* __restrict__ rbuf) {
  T* k = state + S_NARR * ncols;
  if (k[K_DONE] != T(0)) return;
  CgWs<T> w = cg_ws<T>(ws);
  const int tid = threadIdx.x;
  const int cl = tid % CP, rl = tid / CP, rpb = kCgBlock / CP;
  const int ncb = (ncols + CP - 1) / CP;
  const T eps = k[K_EPS];
  T acc[kCgMaxCB], alpha[kCgMaxCB];
#pragma unroll
  for (int cb = 0; cb < kCgMaxCB; ++cb) {
    acc[cb] = T(0);
    const int c = cb * CP + cl;
    alpha[cb] = (cb < ncb && c < ncols) ? cg_alpha_of<T>(state, ncols, c, eps) : T(0);
  }
  for (int64_t row = (int64_t)blockIdx.x * rpb + rl; row < n; row += (int64_t)gridDim.x * rpb) {
#pragma unroll
    for (int cb = 0; cb < kCgMaxCB; ++cb) {
      const int c = cb * CP + cl;
      if (cb < ncb && c < ncols) {
        const int64_t o = row * ld + c;
        const T a = alpha[cb];
        const T rn = fma(-a, v[o], r[o]);
        r[o] = rn;
        x[o] = fma(a, p[o], x[o]);
        acc[cb] = fma(rn, rn, acc[cb]);
      }
    }
  }
  block_col_reduce<T>(acc, ncb, CP, ncols, w.partials);
  if (last_block_ticket(w.counter)) {
    __shared__ T tot[kCgMaxCols];
    last_block_reduce<T>(w.partials, ncols, tot);
    if (rbuf) cg_export_tot<T>(tot, rbuf, ncols);
    else cg_finish_update<T>(state, tot, ncols, hist, max_hist);
  }
}

// ---- vectorised fast paths (ld a power of two <= 128: every thread keeps the same columns across the grid stride) -----
template <typename T> struct V16;
template <> struct V16<float> { using type = float4; static constexpr int N = 4; };
template <> struct V16<double> { using type = double2; static constexpr int N = 2; };

template <typename T>
__device__ __forceinline__ void v16_unpack(const typename V16<T>::type& v, T* o);
template <> __device__ __forceinline__ void v16_unpack<float>(const float4& v, float* o) { o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w; }
template <> __device__ __forceinline__ void v16_unpack<double>(const double2& v, double* o) { o[0] = v.x; o[1] = v.y; }
template <typename T>
__device__ __forceinline__ typename V16<T>::type v16_pack(const T* o);
template <> __device__ __forceinline__ float4 v16_pack<float>(const float* o) { return make_float4(o[0], o[1], o[2], o[3]); }
template <> __device__ __forceinline__ double2 v16_pack<double>(const double* o) { return make_double2(o[0], o[1]); }

template <typename T>
__global__ void __launch_bounds__(kCgBlock)
cg_update_vec_kernel(T* __restrict__ x, T* __restrict__ r, const T* __restrict__ p, const T* __restrict__ v, int ld,
                     int64_t n, int ncols, T* __restrict__ state, T* __restrict__ hist, int max_hist, void* ws,
                     T* __restrict__ rbuf) {
  using V = typename V16<T>::type;
  constexpr int N = V16<T>::N;
  T* k = state + S_NARR * ncols;
  if (k[K_DONE] != T(0)) return;
  CgWs<T> w = cg_ws<T>(ws);
  __shared__ T al[kCgMaxCols];
  const int tid = threadIdx.x;
  const T eps = k[K_EPS];
  for (int c = tid; c < ld; c += kCgBlock) al[c] = c < ncols ? cg_alpha_of<T>(state, ncols, c, eps) : T(0);
  __syncthreads();
  const int c0 = (tid * N) % ld;          // this thread's columns: constant because (blockDim * N) % ld == 0
  T a[N], acc[N];
#pragma unroll
  for (int u = 0; u < N; ++u) { a[u] = al[c0 + u]; acc[u] = T(0); }
  const int64_t total = n * (int64_t)ld / N;
  const int64_t stride = (int64_t)gridDim.x * kCgBlock;
  V* x4 = reinterpret_cast<V*>(x);
  V* r4 = reinterpret_cast<V*>(r);
  const V* p4 = reinterpret_cast<const V*>(p);
  const V* v4 = reinterpret_cast<const V*>(v);
#pragma unroll 2
  for (int64_t e = (int64_t)blockIdx.x * kCgBlock + tid; e < total; e += stride) {
    T pv[N], vv[N], xv[N], rv[N];
    v16_unpack<T>(p4[e], pv); v16_unpack<T>(v4[e], vv); v16_unpack<T>(x4[e], xv); v16_unpack<T>(r4[e], rv);
#pragma unroll
    for (int u = 0; u < N; ++u) {
      rv[u] = fma(-a[u], vv[u], rv[u]);
      xv[u] = fma(a[u], pv[u], xv[u]);
      acc[u] = fma(rv[u], rv[u], acc[u]);
    }
    r4[e] = v16_pack<T>(rv);
    x4[e] = v16_pack<T>(xv);
  }
  vec_block_partials<T, N>(acc, ld, ncols, w.partials + (int64_t)blockIdx.x * ncols);
  if (last_block_ticket_writers(w.counter, tid < ncols)) {
    __shared__ T tot[kCgMaxCols];
    last_block_reduce<T>(w.partials, ncols, tot);
    if (rbuf) cg_export_tot<T>(tot, rbuf, ncols);
    else cg_finish_update<T>(state, tot, ncols, hist, max_hist);
  }
}

// ---- one-block scalar kernels of the multi-GPU path: finish the bookkeeping from all-reduced column sums ----------------
template <typename T>
__global__ void __launch_bounds__(kCgBlock)
cg_scalars_kernel(T* __restrict__ state, const T* __restrict__ rbuf, int ncols, int what, T tol, T eps, T stop, int max_iter,
                  int n_tridiag_iter, T* __restrict__ hist, int max_hist) {
  __shared__ T tot[kCgMaxCols];
  if (what == 2 && state[S_NARR * ncols + K_DONE] != T(0)) return;
  for (int c = threadIdx.x; c < ncols; c += kCgBlock) tot[c] = rbuf[c];
  __syncthreads();
  if (what == 0) cg_finish_norm<T>(state, tot, ncols, eps);
  else if (what == 1) cg_finish_init<T>(state, tot, ncols, tol, eps, stop, max_iter, n_tridiag_iter);
  else cg_finish_update<T>(state, tot, ncols, hist, max_hist);
}

template <typename T>
__global__ void __launch_bounds__(kCgBlock)
cg_pupdate_vec_kernel(T* __restrict__ p, const T* __restrict__ r, int ld, int64_t n, int ncols, const T* __restrict__ state) {
  using V = typename V16<T>::type;
  constexpr int N = V16<T>::N;
  const int c0 = (threadIdx.x * N) % ld;
  T b[N];
#pragma unroll
  for (int u = 0; u < N; ++u) b[u] = (c0 + u) < ncols ? state[S_BETA * ncols + c0 + u] : T(0);
  const int64_t total = n * (int64_t)ld / N;
  const int64_t stride = (int64_t)gridDim.x * kCgBlock;
  V* p4 = reinterpret_cast<V*>(p);
  const V* r4 = reinterpret_cast<const V*>(r);
#pragma unroll 4
  for (int64_t e = (int64_t)blockIdx.x * kCgBlock + threadIdx.x; e < total; e += stride) {
    T pv[N], rv[N];
    v16_unpack<T>(p4[e], pv); v16_unpack<T>(r4[e], rv);
#pragma unroll
    for (int u = 0; u < N; ++u) pv[u] = fma(b[u], pv[u], rv[u]);
    p4[e] = v16_pack<T>(pv);
  }
}

template <typename T>
static inline bool cg_vec_ok(int64_t ld, const void* a, const void* b, const void* c, const void* d) {
  constexpr int N = V16<T>::N;
  if (ld < N || ld > kCgMaxCols || (ld & (ld - 1)) != 0) return false;
  auto al = [](const void* q) { return q == nullptr || (((uintptr_t)q) % 16) == 0; };
  return al(a) && al(b) && al(c) && al(d);
}

template <typename T>
__global__ void __launch_bounds__(kCgBlock)
cg_pupdate_kernel(T* __restrict__ p, const T* __restrict__ r, int64_t ld, int64_t n, int ncols,
                  const T* __restrict__ state) {
  // runs even on the final iteration (as the published loop does); p is not used afterwards
  const int64_t total = n * ld;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(e % ld);
    if (c < ncols) p[e] = fma(state[S_BETA * ncols + c], p[e], r[e]);
  }
}

template <typename T>
__global__ void __launch_bounds__(kCgBlock)
cg_finalize_kernel(const T* __restrict__ x, int64_t ld, T* __restrict__ out, int64_t ldo, int64_t n, int ncols,
                   const T* __restrict__ state) {
  const int64_t total = n * ncols;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = e / ncols;
    const int c = (int)(e - row * ncols);
    out[row * ldo + c] = x[row * ld + c] * state[S_RHSNORM * ncols + c];
  }
}

template <typename T>
static int cg_check(int64_t n, int ncols, int64_t ld) {
  MGP_CHECK_ARG(n > 0 && ncols > 0 && ncols <= kCgMaxCols, "cg: need 0 < ncols <= %d (got %d), n > 0", kCgMaxCols, ncols);
  MGP_CHECK_ARG(ld >= ncols, "cg: ld %lld < ncols %d", (long long)ld, ncols);
  return MGP_OK;
}

template <typename T>
static int cg_init(const T* b, int64_t ldb, T* x, T* r, T* p, int64_t ld, int64_t n, int ncols, T tol, T eps, T stop,
                   int max_iter, int n_tridiag_iter, T* state, void* ws, cudaStream_t st, int phase = 0, T* rbuf = nullptr) {
  int rc = cg_check<T>(n, ncols, ld);
  if (rc) return rc;
  MGP_CHECK_ARG(b && x && r && p && state && ws && ldb >= ncols, "cg_init: bad arguments");
  const int CP = col_lanes(ncols);
  const int grid = cg_grid(n, kCgBlock / CP);
  // phase 0: both passes (single GPU); phase 1 / 2: one pass each, column sums exported to rbuf (multi-GPU)
  if (phase == 0 || phase == 1) {
    cg_norm_kernel<T><<<grid, kCgBlock, 0, st>>>(b, ldb, n, ncols, CP, eps, state, ws, rbuf);
    MGP_LAUNCH_CHECK();
  }
  if (phase == 0 || phase == 2) {
    cg_init_kernel<T><<<grid, kCgBlock, 0, st>>>(b, ldb, x, r, p, ld, n, ncols, CP, tol, eps, stop, max_iter, n_tridiag_iter, state, ws, rbuf);
    MGP_LAUNCH_CHECK();
  }
  return MGP_OK;
}

template <typename T>
static int cg_alpha(const T* p, const T* v, int64_t ld, int64_t n, int ncols, int have_pap, T* state, void* ws,
                    cudaStream_t st) {
  int rc = cg_check<T>(n, ncols, ld);
  if (rc) return rc;
  if (have_pap) return MGP_OK;  // alpha itself is derived inside cg_update from (rz, pAp, masks)
  MGP_CHECK_ARG(p && v && state && ws, "cg_alpha: null pointer");
  const int CP = col_lanes(ncols);
  cg_pap_kernel<T><<<cg_grid(n, kCgBlock / CP), kCgBlock, 0, st>>>(p, v, ld, n, ncols, CP, state, ws);
  MGP_LAUNCH_CHECK();
  return MGP_OK;
}

template <typename T>
static int cg_update(T* x, T* r, const T* p, const T* v, int64_t ld, int64_t n, int ncols, T* state, T* hist,
                     int max_hist, void* ws, cudaStream_t st, T* rbuf = nullptr) {
  int rc = cg_check<T>(n, ncols, ld);
  if (rc) return rc;
  MGP_CHECK_ARG(x && r && p && v && state && ws, "cg_update: null pointer");
  if (cg_vec_ok<T>(ld, x, r, p, v)) {
    int64_t g = ceil_div(n * ld / V16<T>::N, (int64_t)kCgBlock * 4);
    const int64_t cap = (int64_t)kNumSMs * 8;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    cg_update_vec_kernel<T><<<(unsigned)g, kCgBlock, 0, st>>>(x, r, p, v, (int)ld, n, ncols, state, hist, max_hist, ws, rbuf);
    MGP_LAUNCH_CHECK();
    return MGP_OK;
  }
  const int CP = col_lanes(ncols);
  cg_update_kernel<T><<<cg_grid(n, kCgBlock / CP), kCgBlock, 0, st>>>(x, r, p, v, ld, n, ncols, CP, state, hist, max_hist, ws, rbuf);
  MGP_LAUNCH_CHECK();
  return MGP_OK;
}


// ---- split update (8 instead of 9 vector passes per iteration) ---------------------------------------------------------
//   cg_rupdate : r -= alpha v ; rz' = |r|^2 ; scalars            (reads r, v; writes r)
//   cg_pxupdate: x += alpha p ; p = r + beta p                   (reads x, p, r; writes x, p)
// x only needs alpha_k and p_k, both still available when p is rewritten, so its update rides on the pass that reads p
// anyway.  alpha_k is kept in state[S_ALPHA] by the scalar step; K_XPEND says "the x update of the last executed
// iteration is pending", so the pass still runs once after the iteration that set the done flag and never again.
template <typename T>
__global__ void __launch_bounds__(kCgBlock)
cg_rupdate_vec_kernel(T* __restrict__ r, const T* __restrict__ v, int ld, int64_t n, int ncols, T* __restrict__ state,
                      T* __restrict__ hist, int max_hist, void* ws, T* __restrict__ rbuf) {
  using V = typename V16<T>::type;
  constexpr int N = V16<T>::N;
  pdl_wait();
  pdl_launch_dependents();
  T* k = state + S_NARR * ncols;
  if (k[K_DONE] != T(0)) {
    if (blockIdx.x == 0 && threadIdx.x == 0) k[K_XPEND] = T(0);
    return;
  }
  CgWs<T> w = cg_ws<T>(ws);
  __shared__ T al[kCgMaxCols];
  const int tid = threadIdx.x;
  const T eps = k[K_EPS];
  for (int c = tid; c < ld; c += kCgBlock) al[c] = c < ncols ? cg_alpha_of<T>(state, ncols, c, eps) : T(0);
  __syncthreads();
  // reversed walk (below): element total-1-q of thread q = tid (mod ld/N) has the MIRRORED columns
  const int c0 = (ld - ((tid + 1) * N) % ld) % ld;
  T a[N], acc[N];
#pragma unroll
  for (int u = 0; u < N; ++u) { a[u] = al[c0 + u]; acc[u] = T(0); }
  const int64_t total = n * (int64_t)ld / N;
  const int64_t stride = (int64_t)gridDim.x * kCgBlock;
  V* r4 = reinterpret_cast<V*>(r);
  const V* v4 = reinterpret_cast<const V*>(v);
  // Walks the vectors from the END: the matvec that produced v wrote its rows in ascending tile order, so the tail of v is
  // what is still in L2; the p / x pass that follows walks ascending and meets the rows of r this pass wrote last.
  // (total * N) % ld == 0 and (kCgBlock * N) % ld == 0, so element total-1-q has the columns of a thread whose index is the
  // mirror image: the per-thread column map is re-derived for the reversed index.
#pragma unroll 4
  for (int64_t q = (int64_t)blockIdx.x * kCgBlock + tid; q < total; q += stride) {
    const int64_t e = total - 1 - q;
    T vv[N], rv[N];
    v16_unpack<T>(__ldcs(v4 + e), vv); v16_unpack<T>(r4[e], rv);
#pragma unroll
    for (int u = 0; u < N; ++u) {
      rv[u] = fma(-a[u], vv[u], rv[u]);
      acc[u] = fma(rv[u], rv[u], acc[u]);
    }
    r4[e] = v16_pack<T>(rv);
  }
  vec_block_partials<T, N>(acc, ld, ncols, w.partials + (int64_t)blockIdx.x * ncols, c0);
  if (last_block_ticket_writers(w.counter, tid < ncols)) {     // only the threads that wrote the block's partials fence
    __shared__ T tot[kCgMaxCols];
    last_block_reduce<T>(w.partials, ncols, tot);
    if (rbuf) cg_export_tot<T>(tot, rbuf, ncols);
    else cg_finish_update<T>(state, tot, ncols, hist, max_hist);
  }
}

template <typename T>
__global__ void __launch_bounds__(kCgBlock)
cg_pxupdate_vec_kernel(T* __restrict__ x, T* __restrict__ p, const T* __restrict__ r, int ld, int64_t n, int ncols,
                       const T* __restrict__ state) {
  using V = typename V16<T>::type;
  constexpr int N = V16<T>::N;
  pdl_wait();
  pdl_launch_dependents();
  if (state[S_NARR * ncols + K_XPEND] == T(0)) return;
  const int c0 = (threadIdx.x * N) % ld;
  T a[N], b[N];
#pragma unroll
  for (int u = 0; u < N; ++u) {
    a[u] = (c0 + u) < ncols ? state[S_ALPHA * ncols + c0 + u] : T(0);
    b[u] = (c0 + u) < ncols ? state[S_BETA * ncols + c0 + u] : T(0);
  }
  const int64_t total = n * (int64_t)ld / N;
  const int64_t stride = (int64_t)gridDim.x * kCgBlock;
  V* x4 = reinterpret_cast<V*>(x);
  V* p4 = reinterpret_cast<V*>(p);
  const V* r4 = reinterpret_cast<const V*>(r);
#pragma unroll 4
  for (int64_t e = (int64_t)blockIdx.x * kCgBlock + threadIdx.x; e < total; e += stride) {
    T pv[N], rv[N], xv[N];
    v16_unpack<T>(p4[e], pv); v16_unpack<T>(r4[e], rv); v16_unpack<T>(__ldcs(x4 + e), xv);
#pragma unroll
    for (int u = 0; u < N; ++u) {
      xv[u] = fma(a[u], pv[u], xv[u]);
      pv[u] = fma(b[u], pv[u], rv[u]);
    }
    __stcs(x4 + e, v16_pack<T>(xv));
    p4[e] = v16_pack<T>(pv);
  }
}

// generic layouts (any ld): scalar grid-stride versions of the same two passes
template <typename T>
__global__ void __launch_bounds__(kCgBlock)
cg_rupdate_kernel(T* __restrict__ r, const T* __restrict__ v, int64_t ld, int64_t n, int ncols, int CP, T* __restrict__ state,
                  T* __restrict__ hist, int max_hist, void* ws, T* __restrict__ rbuf) {
  T* k = state + S_NARR * ncols;
  if (k[K_DONE] != T(0)) {
    if (blockIdx.x == 0 && threadIdx.x == 0) k[K_XPEND] = T(0);
    return;
  }
  CgWs<T> w = cg_ws<T>(ws);
  const int tid = threadIdx.x;
  const int cl = tid % CP, rl = tid / CP, rpb = kCgBlock / CP;
  const int ncb = (ncols + CP - 1) / CP;
  const T eps = k[K_EPS];
  T acc[kCgMaxCB], alpha[kCgMaxCB];
#pragma unroll
  for (int cb = 0; cb < kCgMaxCB; ++cb) {
    acc[cb] = T(0);
    const int c = cb * CP + cl;
    alpha[cb] = (cb < ncb && c < ncols) ? cg_alpha_of<T>(state, ncols, c, eps) : T(0);
  }
  for (int64_t row = (int64_t)blockIdx.x * rpb + rl; row < n; row += (int64_t)gridDim.x * rpb) {
#pragma unroll
    for (int cb = 0; cb < kCgMaxCB; ++cb) {
      const int c = cb * CP + cl;
      if (cb < ncb && c < ncols) {
        const int64_t o = row * ld + c;
        const T rn = fma(-alpha[cb], v[o], r[o]);
        r[o] = rn;
        acc[cb] = fma(rn, rn, acc[cb]);
      }
    }
  }
  block_col_reduce<T>(acc, ncb, CP, ncols, w.partials);
  if (last_block_ticket(w.counter)) {
    __shared__ T tot[kCgMaxCols];
    last_block_reduce<T>(w.partials, ncols, tot);
    if (rbuf) cg_export_tot<T>(tot, rbuf, ncols);
    else cg_finish_update<T>(state, tot, ncols, hist, max_hist);
  }
}

template <typename T>
__global__ void __launch_bounds__(kCgBlock)
cg_pxupdate_kernel(T* __restrict__ x, T* __restrict__ p, const T* __restrict__ r, int64_t ld, int64_t n, int ncols,
                   const T* __restrict__ state) {
  if (state[S_NARR * ncols + K_XPEND] == T(0)) return;
  const int64_t total = n * ld;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(e % ld);
    if (c < ncols) {
      const T pv = p[e];
      x[e] = fma(state[S_ALPHA * ncols + c], pv, x[e]);
      p[e] = fma(state[S_BETA * ncols + c], pv, r[e]);
    }
  }
}

template <typename T>
static int cg_rupdate(T* r, const T* v, int64_t ld, int64_t n, int ncols, T* state, T* hist, int max_hist, void* ws,
                      cudaStream_t st, T* rbuf = nullptr) {
  int rc = cg_check<T>(n, ncols, ld);
  if (rc) return rc;
  MGP_CHECK_ARG(r && v && state && ws, "cg_rupdate: null pointer");
  if (cg_vec_ok<T>(ld, r, v, nullptr, nullptr)) {
    int64_t g = ceil_div(n * ld / V16<T>::N, (int64_t)kCgBlock * 4);
    const int64_t cap = (int64_t)kNumSMs * 8;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    MGP_CUDA(launch_pdl(cg_rupdate_vec_kernel<T>, dim3((unsigned)g), dim3(kCgBlock), 0, st, r, v, (int)ld, n, ncols, state, hist, max_hist, ws, rbuf));
    MGP_LAUNCH_CHECK();
    return MGP_OK;
  }
  const int CP = col_lanes(ncols);
  cg_rupdate_kernel<T><<<cg_grid(n, kCgBlock / CP), kCgBlock, 0, st>>>(r, v, ld, n, ncols, CP, state, hist, max_hist, ws, rbuf);
  MGP_LAUNCH_CHECK();
  return MGP_OK;
}

template <typename T>
static int cg_pxupdate(T* x, T* p, const T* r, int64_t ld, int64_t n, int ncols, const T* state, cudaStream_t st) {
  MGP_CHECK_ARG(x && p && r && state && n > 0 && ncols > 0 && ld >= ncols, "cg_pxupdate: bad arguments");
  const int64_t total = n * ld;
  if (cg_vec_ok<T>(ld, x, p, r, nullptr)) {
    int64_t gv = ceil_div(total / V16<T>::N, (int64_t)kCgBlock * 4);
    if (gv > kNumSMs * 8) gv = kNumSMs * 8;
    if (gv < 1) gv = 1;
    MGP_CUDA(launch_pdl(cg_pxupdate_vec_kernel<T>, dim3((unsigned)gv), dim3(kCgBlock), 0, st, x, p, r, (int)ld, n, ncols, state));
    MGP_LAUNCH_CHECK();
    return MGP_OK;
  }
  int64_t g = ceil_div(total, (int64_t)kCgBlock);
  if (g > kNumSMs * 16) g = kNumSMs * 16;
  cg_pxupdate_kernel<T><<<(unsigned)g, kCgBlock, 0, st>>>(x, p, r, ld, n, ncols, state);
  MGP_LAUNCH_CHECK();
  return MGP_OK;
}


// ---- multi-GPU, peer-memory path: all-reduce of the column sums + scalar step in ONE one-block kernel ---------------------
// NCCL's all-reduce of 16 floats costs ~37 us per call on the 8 x B200 box (measured, profiles/), more than a rank's whole
// SpMM at N = 1M / 8.  Here every rank stores its C partial sums straight into slot [my_rank] of every peer's `red` buffer
// (NVLink P2P stores), publishes an epoch number in every peer's flag array with a system-scope release store, spins until
// all peers' flags carry the epoch, and adds the world partials in rank order -- so every rank computes bit-identical
// totals -- before running the same scalar step as the single-GPU kernels.  Two alternating `red` buffers are enough: a
// peer can only write epoch e+2 after it has seen my flag of epoch e+1, which I set after reading epoch e.
//   red_ptrs[r]  -> rank r's buffer  T[2][world][kCgMaxCols]           (peer-mapped addresses, device array)
//   flag_ptrs[r] -> rank r's flags   unsigned[world]                    (one flag per source rank)
//   epoch        -> this rank's private counter (device memory), advanced by the kernel: CUDA-graph replay safe
// what: 0 norm, 1 init, 2 update (as cg_scalars_kernel), 3 = store the totals as p^T A p.
// all peers' flags[src] >= epoch (wrap-safe); bounded so a lost peer traps instead of hanging the GPU
__device__ __forceinline__ void wait_flags(const unsigned int* my_flags, int world, unsigned int epoch) {
  if ((int)threadIdx.x < world) {
    unsigned long long spins = 0;
    while ((int)(ld_acquire_sys(my_flags + threadIdx.x) - epoch) < 0) {
      if (++spins > (1ull << 25)) __trap();     // ~10 s: a peer died or the call sequences diverged
    }
  }
  __syncthreads();
}

template <typename T>
__global__ void __launch_bounds__(kCgBlock)
cg_peer_scalars_kernel(T* __restrict__ state, const T* __restrict__ rbuf, int ncols, int what, T tol, T eps, T stop,
                       int max_iter, int n_tridiag_iter, T* __restrict__ hist, int max_hist, T* const* __restrict__ red_ptrs,
                       unsigned int* const* __restrict__ flag_ptrs, unsigned int* __restrict__ epoch_ctr, int rank, int world) {
  __shared__ T tot[kCgMaxCols];
  const int tid = threadIdx.x;
  const unsigned int epoch = *epoch_ctr + 1u;
  const int buf = (int)(epoch & 1u);
  // 1. my partials -> every rank's red[buf][rank][:]
  for (int i = tid; i < world * ncols; i += kCgBlock) {
    const int r = i / ncols, c = i - r * ncols;
    red_ptrs[r][((size_t)buf * world + rank) * kCgMaxCols + c] = rbuf[c];
  }
  __threadfence_system();
  __syncthreads();
  // 2. publish, 3. wait for everyone
  if (tid < world) st_release_sys(flag_ptrs[tid] + rank, epoch);
  wait_flags(flag_ptrs[rank], world, epoch);
  // 4. totals in rank order
  const T* mine = red_ptrs[rank] + (size_t)buf * world * kCgMaxCols;
  for (int c = tid; c < ncols; c += kCgBlock) {
    T s = T(0);
    for (int r = 0; r < world; ++r) s += __ldcv(mine + (size_t)r * kCgMaxCols + c);
    tot[c] = s;
  }
  __syncthreads();
  if (tid == 0) *epoch_ctr = epoch;
  if (what == 3) {
    for (int c = tid; c < ncols; c += kCgBlock) state[S_PAP * ncols + c] = tot[c];
  } else if (what == 2) {
    if (state[S_NARR * ncols + K_DONE] == T(0)) cg_finish_update<T>(state, tot, ncols, hist, max_hist);
  } else if (what == 0) {
    cg_finish_norm<T>(state, tot, ncols, eps);
  } else {
    cg_finish_init<T>(state, tot, ncols, tol, eps, stop, max_iter, n_tridiag_iter);
  }
}

// cross-GPU barrier (one block): "everything this rank enqueued before is visible to its peers, and theirs to me"
__global__ void __launch_bounds__(32)
peer_barrier_kernel(unsigned int* const* __restrict__ flag_ptrs, unsigned int* __restrict__ epoch_ctr, int rank, int world) {
  const unsigned int epoch = *epoch_ctr + 1u;
  __threadfence_system();
  if ((int)threadIdx.x < world) st_release_sys(flag_ptrs[threadIdx.x] + rank, epoch);
  wait_flags(flag_ptrs[rank], world, epoch);
  if (threadIdx.x == 0) *epoch_ctr = epoch;
}


// ---- multi-GPU, fully fused iteration tail: [rupdate + both all-reduces + scalar step + pxupdate] in TWO kernels ----------
// Sync points are identified by the CG iteration number itself (epoch = K_ITER + 1, identical on every rank), so a captured
// CUDA graph can be replayed and nothing has to be counted on the side; the flags are zeroed once per solve.
//   flag2 / flag3 : uint32[world] per rank (peer-mapped) -- "p^T A p partials of rank r have landed" / "|r|^2 partials ..."
//   red_ptrs[r]   : T[2 kinds][2 buffers][world][kCgMaxCols] of rank r
// cg_peer_rupdate : block 0 ships this rank's p^T A p partials (dot epilogue of the last SpMM launch) to every peer; every
//                   block waits for all ranks' partials, adds them in rank order, forms alpha with the published safe
//                   division, updates r and accumulates |r|^2; the last block to finish ships the |r|^2 partials.
// cg_peer_pxupdate: every block waits for all ranks' |r|^2 partials, forms beta, applies x += alpha p, p = r + beta p; the
//                   last block to finish runs the scalar step (norms, flags, history, iteration counter).
template <typename T>
__device__ __forceinline__ T* peer_red_slot(T* base, int kind, int buf, int world, int src) {
  return base + (((size_t)kind * 2 + buf) * world + src) * kCgMaxCols;
}

template <typename T>
__global__ void __launch_bounds__(kCgBlock)
cg_peer_rupdate_kernel(T* __restrict__ r, const T* __restrict__ v, int ld, int64_t n, int ncols, T* __restrict__ state,
                       const T* __restrict__ pap_local, void* ws, T* const* __restrict__ red_ptrs,
                       unsigned int* const* __restrict__ flag2_ptrs, unsigned int* const* __restrict__ flag3_ptrs, int rank,
                       int world) {
  using V = typename V16<T>::type;
  constexpr int N = V16<T>::N;
  T* k = state + S_NARR * ncols;
  if (k[K_DONE] != T(0)) return;
  CgWs<T> w = cg_ws<T>(ws);
  __shared__ T al[kCgMaxCols];
  const int tid = threadIdx.x;
  const unsigned int epoch = (unsigned int)k[K_ITER] + 1u;
  const int buf = (int)(epoch & 1u);
  if (blockIdx.x == 0) {                                   // ship my p^T A p partials
    for (int i = tid; i < world * ncols; i += kCgBlock) {
      const int dst = i / ncols, c = i - dst * ncols;
      peer_red_slot<T>(red_ptrs[dst], 0, buf, world, rank)[c] = pap_local[c];
    }
    __threadfence_system();
    __syncthreads();
    if (tid < world) st_release_sys(flag2_ptrs[tid] + rank, epoch);
  }
  wait_flags(flag2_ptrs[rank], world, epoch);
  const T eps = k[K_EPS];
  for (int c = tid; c < ld; c += kCgBlock) {
    T a = T(0);
    if (c < ncols) {
      T pap = T(0);
      for (int src = 0; src < world; ++src) pap += __ldcv(peer_red_slot<T>(red_ptrs[rank], 0, buf, world, src) + c);
      const T rz = state[S_RZ * ncols + c];
      a = (pap < eps) ? T(0) : rz / pap;
      if (state[S_CONV * ncols + c] != T(0)) a = T(0);
      if (blockIdx.x == 0) { state[S_PAP * ncols + c] = pap; state[S_ALPHA * ncols + c] = a; }   // read by nobody in this kernel
    }
    al[c] = a;
  }
  __syncthreads();
  const int c0 = (tid * N) % ld;
  T a[N], acc[N];
#pragma unroll
  for (int u = 0; u < N; ++u) { a[u] = al[c0 + u]; acc[u] = T(0); }
  const int64_t total = n * (int64_t)ld / N;
  const int64_t stride = (int64_t)gridDim.x * kCgBlock;
  V* r4 = reinterpret_cast<V*>(r);
  const V* v4 = reinterpret_cast<const V*>(v);
#pragma unroll 4
  for (int64_t e = (int64_t)blockIdx.x * kCgBlock + tid; e < total; e += stride) {
    T vv[N], rv[N];
    v16_unpack<T>(v4[e], vv); v16_unpack<T>(r4[e], rv);
#pragma unroll
    for (int u = 0; u < N; ++u) {
      rv[u] = fma(-a[u], vv[u], rv[u]);
      acc[u] = fma(rv[u], rv[u], acc[u]);
    }
    r4[e] = v16_pack<T>(rv);
  }
  vec_block_partials<T, N>(acc, ld, ncols, w.partials + (int64_t)blockIdx.x * ncols);
  if (last_block_ticket_writers(w.counter, tid < ncols)) {
    __shared__ T tot[kCgMaxCols];
    last_block_reduce<T>(w.partials, ncols, tot);
    for (int i = tid; i < world * ncols; i += kCgBlock) {   // ship my |r|^2 partials
      const int dst = i / ncols, c = i - dst * ncols;
      peer_red_slot<T>(red_ptrs[dst], 1, buf, world, rank)[c] = tot[c];
    }
    __threadfence_system();
    __syncthreads();
    if (tid < world) st_release_sys(flag3_ptrs[tid] + rank, epoch);
    if (tid == 0) k[K_XPEND] = T(1);
  }
}

template <typename T>
__global__ void __launch_bounds__(kCgBlock)
cg_peer_pxupdate_kernel(T* __restrict__ x, T* __restrict__ p, const T* __restrict__ r, int ld, int64_t n, int ncols,
                        T* __restrict__ state, T* __restrict__ hist, int max_hist, void* ws, T* const* __restrict__ red_ptrs,
                        unsigned int* const* __restrict__ flag3_ptrs, int rank, int world) {
  using V = typename V16<T>::type;
  constexpr int N = V16<T>::N;
  T* k = state + S_NARR * ncols;
  if (k[K_XPEND] == T(0)) return;
  __shared__ T al[kCgMaxCols], be[kCgMaxCols], tot[kCgMaxCols];
  const int tid = threadIdx.x;
  const unsigned int epoch = (unsigned int)k[K_ITER] + 1u;
  const int buf = (int)(epoch & 1u);
  wait_flags(flag3_ptrs[rank], world, epoch);
  const T eps = k[K_EPS];
  for (int c = tid; c < ld; c += kCgBlock) {
    T a = T(0), b = T(0);
    if (c < ncols) {
      T rz_new = T(0);
      for (int src = 0; src < world; ++src) rz_new += __ldcv(peer_red_slot<T>(red_ptrs[rank], 1, buf, world, src) + c);
      const T rz_old = state[S_RZ * ncols + c];
      b = (rz_old < eps) ? T(0) : rz_new / rz_old;
      a = state[S_ALPHA * ncols + c];
      tot[c] = rz_new;
    }
    al[c] = a; be[c] = b;
  }
  __syncthreads();
  const int c0 = (tid * N) % ld;
  T a[N], b[N];
#pragma unroll
  for (int u = 0; u < N; ++u) { a[u] = al[c0 + u]; b[u] = be[c0 + u]; }
  const int64_t total = n * (int64_t)ld / N;
  const int64_t stride = (int64_t)gridDim.x * kCgBlock;
  V* x4 = reinterpret_cast<V*>(x);
  V* p4 = reinterpret_cast<V*>(p);
  const V* r4 = reinterpret_cast<const V*>(r);
#pragma unroll 4
  for (int64_t e = (int64_t)blockIdx.x * kCgBlock + tid; e < total; e += stride) {
    T pv[N], rv[N], xv[N];
    v16_unpack<T>(p4[e], pv); v16_unpack<T>(r4[e], rv); v16_unpack<T>(x4[e], xv);
#pragma unroll
    for (int u = 0; u < N; ++u) {
      xv[u] = fma(a[u], pv[u], xv[u]);
      pv[u] = fma(b[u], pv[u], rv[u]);
    }
    x4[e] = v16_pack<T>(xv);
    p4[e] = v16_pack<T>(pv);
  }
  // every block has read the old scalars: the last one to finish advances them (alpha via cg_alpha_of == S_ALPHA)
  CgWs<T> w = cg_ws<T>(ws);
  if (last_block_ticket(w.counter + 4)) {
    cg_finish_update<T>(state, tot, ncols, hist, max_hist);
    __syncthreads();
    if (tid == 0) k[K_XPEND] = T(0);
  }
}


// ---- multi-GPU, single-reduction iteration (Chronopoulos & Gear): ONE vector kernel, ONE all-reduce per iteration --------------
// Standard CG needs p^T A p before the r update and |r|^2 before the p update: two cross-GPU reductions per iteration, each
// a sync point (~4 us of NVLink round trip + a launch).  The Chronopoulos-Gear recurrences produce the same iterates from
//     w = A r,   gamma = r.r,   delta = r.w        (both dots known as soon as the matvec is done -> one reduction)
//     beta_k  = gamma_k / gamma_{k-1}              (beta_0 = 0)
//     alpha_k = gamma_k / (delta_k - beta_k gamma_k / alpha_{k-1})          (= gamma_k / p_k^T A p_k in exact arithmetic)
//     p = r + beta p ;  s = w + beta s (= A p) ;  x += alpha p ;  r -= alpha s
// so an iteration is [SpMM x nu with r as the source] + this kernel.  The last SpMM launch ships this rank's (delta, gamma)
// partials to every peer when its last block finishes (mgp_lap_spmm_wi_ex: red_ptrs / red_flags); every block here waits
// for all ranks' partials, adds them in rank order (bit-identical scalars on every rank), applies the four updates and
// accumulates |r_new|^2; the last block to finish stores that local sum for the next shipment, advances the scalar state
// (same masks / stopping rules / history as cg_finish_update; the residual norm tested is the one ENTERING the iteration,
// so convergence is noticed one matvec later than in the two-reduction form and no update is applied then) and publishes
// "r complete" (epoch + 1) to the peers, whose next SpMM waits for it at its first remote halo row.
// Measured in fp32 on the cfg-C-conditioned torus (30k points, CPU experiment of round 2): same iteration count as the
// two-reduction form, solution error vs fp64 1.2e-5 (3.2e-6 for the standard recurrences).
template <typename T>
__global__ void __launch_bounds__(kCgBlock)
cg_peer_cgstep_kernel(T* __restrict__ x, T* __restrict__ r, T* __restrict__ p, T* __restrict__ s, const T* __restrict__ wv_,
                      int ld, int64_t n, int ncols, T* __restrict__ state, T* __restrict__ hist, int max_hist, void* ws,
                      T* __restrict__ gamma_loc, const T* __restrict__ delta_loc, T* const* __restrict__ red_ptrs,
                      unsigned int* const* __restrict__ dflag_ptrs, unsigned int* const* __restrict__ rflag_ptrs, int rank, int world) {
  using V = typename V16<T>::type;
  constexpr int N = V16<T>::N;
  pdl_wait();
  pdl_launch_dependents();
  T* k = state + S_NARR * ncols;
  if (k[K_DONE] != T(0)) return;
  CgWs<T> w = cg_ws<T>(ws);
  __shared__ T al[kCgMaxCols], be[kCgMaxCols], gam[kCgMaxCols], den[kCgMaxCols], rnm[kCgMaxCols];
  __shared__ T s_mean;
  __shared__ int s_done;
  const int tid = threadIdx.x;
  const int it = (int)k[K_ITER];
  const unsigned int epoch = (unsigned int)it + 1u;
  const int buf = (int)(epoch & 1u);
  if (delta_loc && blockIdx.x == 0) {
    // first half of the all-reduce done here (round-1 placement): ship this rank's (r.w, |r|^2) partials to every peer.  The
    // system-scope fence sees only this block's few stores; shipping from the END of the SpMM launch instead
    // (mgp_wi_ext.red_ptrs) put a fence behind ~10^5 outstanding stores on the critical path (+4 us per iteration at 125k rows).
    for (int i = tid; i < world * ncols; i += kCgBlock) {
      const int dst = i / ncols, c = i - dst * ncols;
      peer_red_slot<T>(red_ptrs[dst], 0, buf, world, rank)[c] = delta_loc[c];
      peer_red_slot<T>(red_ptrs[dst], 1, buf, world, rank)[c] = gamma_loc[c];
    }
    __threadfence_system();
    __syncthreads();
    if (tid < world) st_release_sys(dflag_ptrs[tid] + rank, epoch);
  }
  wait_flags(dflag_ptrs[rank], world, epoch);
  const T eps = k[K_EPS], stop = k[K_STOP];
  T local = T(0);
  for (int c = tid; c < ld; c += kCgBlock) {
    T a = T(0), b = T(0), gm = T(0), dn = T(0), rn = T(0);
    if (c < ncols) {
      T dl = T(0);
      for (int src = 0; src < world; ++src) {
        dl += __ldcv(peer_red_slot<T>(red_ptrs[rank], 0, buf, world, src) + c);
        gm += __ldcv(peer_red_slot<T>(red_ptrs[rank], 1, buf, world, src) + c);
      }
      rn = dev_sqrt<T>(gm);
      if (state[S_RHSZERO * ncols + c] != T(0)) rn = T(0);
      const T g_old = state[S_RZ * ncols + c], a_old = state[S_ALPHA * ncols + c];
      b = (it == 0 || g_old < eps) ? T(0) : gm / g_old;
      dn = (it == 0 || a_old == T(0)) ? dl : dl - b * gm / a_old;
      a = (dn < eps) ? T(0) : gm / dn;
      if (rn < stop) a = T(0);
      local += rn;
    }
    al[c] = a; be[c] = b; gam[c] = gm; den[c] = dn; rnm[c] = rn;
  }
  const T tot = block_sum_t0<T>(local);
  if (tid == 0) {
    const T mean = tot / T(ncols);
    s_mean = mean;
    int done = 0;
    if (it >= 1) {
      const bool tri_pending = (k[K_NTRIMIN] > T(0)) && (T(it - 1) < k[K_NTRIMIN]);
      if (T(it - 1) >= k[K_MINITER] && mean < k[K_TOL] && !tri_pending) done = 1;
      else if (T(it) >= k[K_MAXITER]) done = 2;
    }
    s_done = done;
  }
  __syncthreads();
  if (s_done) {
    // every block reaches the same verdict from the same totals; block 0 records it.  x already is the answer.
    if (blockIdx.x == 0) {
      for (int c = tid; c < ncols; c += kCgBlock) {
        state[S_RESID * ncols + c] = rnm[c];
        state[S_CONV * ncols + c] = (rnm[c] < stop) ? T(1) : T(0);
        if (hist && it >= 1 && it - 1 < max_hist) hist[((int64_t)(it - 1) * 2 + 1) * ncols + c] = be[c];
      }
      __syncthreads();
      if (tid == 0) { k[K_MEAN] = s_mean; __threadfence(); k[K_DONE] = T(s_done); }
    }
    return;
  }
  const int c0 = (tid * N) % ld;
  T a[N], b[N], acc[N];
#pragma unroll
  for (int u = 0; u < N; ++u) { a[u] = al[c0 + u]; b[u] = be[c0 + u]; acc[u] = T(0); }
  // (A software-pipelined form with the first element's loads issued before the flag wait measured SLOWER at 125k rows per
  // rank -- 26 vs 20.5 us -- because it keeps half as many loads in flight; the plain unrolled walk stays.)
  const int64_t total = n * (int64_t)ld / N;
  const int64_t stride = (int64_t)gridDim.x * kCgBlock;
  V* x4 = reinterpret_cast<V*>(x);
  V* r4 = reinterpret_cast<V*>(r);
  V* p4 = reinterpret_cast<V*>(p);
  V* s4 = reinterpret_cast<V*>(s);
  const V* w4 = reinterpret_cast<const V*>(wv_);
#pragma unroll 2
  for (int64_t e = (int64_t)blockIdx.x * kCgBlock + tid; e < total; e += stride) {
    T pv[N], sv[N], rv[N], wv[N], xv[N];
    v16_unpack<T>(p4[e], pv); v16_unpack<T>(s4[e], sv); v16_unpack<T>(r4[e], rv); v16_unpack<T>(w4[e], wv); v16_unpack<T>(x4[e], xv);
#pragma unroll
    for (int u = 0; u < N; ++u) {
      pv[u] = fma(b[u], pv[u], rv[u]);
      sv[u] = fma(b[u], sv[u], wv[u]);
      xv[u] = fma(a[u], pv[u], xv[u]);
      rv[u] = fma(-a[u], sv[u], rv[u]);
      acc[u] = fma(rv[u], rv[u], acc[u]);
    }
    p4[e] = v16_pack<T>(pv); s4[e] = v16_pack<T>(sv); x4[e] = v16_pack<T>(xv); r4[e] = v16_pack<T>(rv);
  }
  vec_block_partials<T, N>(acc, ld, ncols, w.partials + (int64_t)blockIdx.x * ncols);
  // publishing "r complete" from here needs EVERY thread's stores fenced before the ticket; otherwise only the writers of the partials
  if (rflag_ptrs ? last_block_ticket(w.counter) : last_block_ticket_writers(w.counter, tid < ncols)) {
    __shared__ T tot2[kCgMaxCols];
    last_block_reduce<T>(w.partials, ncols, tot2);
    for (int c = tid; c < ncols; c += kCgBlock) {
      gamma_loc[c] = tot2[c];                               // this rank's |r_{k+1}|^2: shipped by the next matvec
      state[S_RZ * ncols + c] = gam[c];
      state[S_PAP * ncols + c] = den[c];
      state[S_ALPHA * ncols + c] = al[c];
      state[S_BETA * ncols + c] = be[c];
      state[S_RESID * ncols + c] = rnm[c];
      state[S_CONV * ncols + c] = (rnm[c] < stop) ? T(1) : T(0);
      if (hist) {
        if (it < max_hist) hist[((int64_t)it * 2 + 0) * ncols + c] = al[c];
        if (it >= 1 && it - 1 < max_hist) hist[((int64_t)(it - 1) * 2 + 1) * ncols + c] = be[c];
      }
    }
    __syncthreads();
    if (tid == 0) { k[K_MEAN] = s_mean; k[K_ITER] = T(it + 1); }
    if (rflag_ptrs) {                  // optional: publish "r_{k+1} complete" from here (else the next SpMM does, at its start)
      __threadfence_system();
      __syncthreads();
      if (tid < world) st_release_sys(rflag_ptrs[tid] + rank, epoch + 1u);
    }
  }
}

// flags[dst][rank] = value on every rank: "everything this rank enqueued before is complete" without waiting for anybody
__global__ void __launch_bounds__(32)
peer_publish_kernel(unsigned int* const* __restrict__ flag_ptrs, unsigned int value, int rank, int world) {
  __threadfence_system();
  if ((int)threadIdx.x < world) st_release_sys(flag_ptrs[threadIdx.x] + rank, value);
}

}  // namespace mgp

using namespace mgp;

extern "C" {

size_t mgp_cg_state_elems(int32_t ncols) { return (size_t)S_NARR * ncols + K_NSCAL; }

size_t mgp_cg_ws_bytes(int64_t n, int32_t ncols) {
  (void)n;
  return 256 + (size_t)kNumSMs * 8 * (size_t)ncols * 8;
}

int mgp_cg_init_f32(const float* b, int64_t ldb, float* x, float* r, float* p, int64_t ld, int64_t n, int32_t ncols,
                    float tolerance, float eps, float stop_updating_after, int32_t max_iter, int32_t n_tridiag_iter,
                    float* state, void* ws, void* stream) {
  return cg_init<float>(b, ldb, x, r, p, ld, n, ncols, tolerance, eps, stop_updating_after, max_iter, n_tridiag_iter, state, ws, (cudaStream_t)stream);
}
int mgp_cg_init_f64(const double* b, int64_t ldb, double* x, double* r, double* p, int64_t ld, int64_t n, int32_t ncols,
                    double tolerance, double eps, double stop_updating_after, int32_t max_iter, int32_t n_tridiag_iter,
                    double* state, void* ws, void* stream) {
  return cg_init<double>(b, ldb, x, r, p, ld, n, ncols, tolerance, eps, stop_updating_after, max_iter, n_tridiag_iter, state, ws, (cudaStream_t)stream);
}
int mgp_cg_alpha_f32(const float* p, const float* v, int64_t ld, int64_t n, int32_t ncols, int32_t have_pap, float* state,
                     void* ws, void* stream) {
  return cg_alpha<float>(p, v, ld, n, ncols, have_pap, state, ws, (cudaStream_t)stream);
}
int mgp_cg_alpha_f64(const double* p, const double* v, int64_t ld, int64_t n, int32_t ncols, int32_t have_pap,
                     double* state, void* ws, void* stream) {
  return cg_alpha<double>(p, v, ld, n, ncols, have_pap, state, ws, (cudaStream_t)stream);
}
int mgp_cg_update_f32(float* x, float* r, const float* p, const float* v, int64_t ld, int64_t n, int32_t ncols,
                      float* state, float* hist, int32_t max_hist, void* ws, void* stream) {
  return cg_update<float>(x, r, p, v, ld, n, ncols, state, hist, max_hist, ws, (cudaStream_t)stream);
}
int mgp_cg_update_f64(double* x, double* r, const double* p, const double* v, int64_t ld, int64_t n, int32_t ncols,
                      double* state, double* hist, int32_t max_hist, void* ws, void* stream) {
  return cg_update<double>(x, r, p, v, ld, n, ncols, state, hist, max_hist, ws, (cudaStream_t)stream);
}
int mgp_cg_rupdate_f32(float* r, const float* v, int64_t ld, int64_t n, int32_t ncols, float* state, float* hist,
                       int32_t max_hist, float* rbuf, void* ws, void* stream) {
  return cg_rupdate<float>(r, v, ld, n, ncols, state, hist, max_hist, ws, (cudaStream_t)stream, rbuf);
}
int mgp_cg_rupdate_f64(double* r, const double* v, int64_t ld, int64_t n, int32_t ncols, double* state, double* hist,
                       int32_t max_hist, double* rbuf, void* ws, void* stream) {
  return cg_rupdate<double>(r, v, ld, n, ncols, state, hist, max_hist, ws, (cudaStream_t)stream, rbuf);
}
int mgp_cg_pxupdate_f32(float* x, float* p, const float* r, int64_t ld, int64_t n, int32_t ncols, const float* state, void* stream) {
  return cg_pxupdate<float>(x, p, r, ld, n, ncols, state, (cudaStream_t)stream);
}
int mgp_cg_pxupdate_f64(double* x, double* p, const double* r, int64_t ld, int64_t n, int32_t ncols, const double* state, void* stream) {
  return cg_pxupdate<double>(x, p, r, ld, n, ncols, state, (cudaStream_t)stream);
}
// ---- multi-GPU split entry points (column sums exported to rbuf, finished by mgp_cg_dist_scalars after the all-reduce) ----
int mgp_cg_dist_norm2_f32(const float* b, int64_t ldb, int64_t n, int32_t ncols, float* state, float* rbuf, void* ws, void* stream) {
  MGP_CHECK_ARG(rbuf != nullptr, "cg_dist_norm2: rbuf is null");
  return cg_init<float>(b, ldb, (float*)b, (float*)b, (float*)b, ncols > ldb ? ldb : ldb, n, ncols, 0.f, 0.f, 0.f, 1, 0, state, ws, (cudaStream_t)stream, 1, rbuf);
}
int mgp_cg_dist_norm2_f64(const double* b, int64_t ldb, int64_t n, int32_t ncols, double* state, double* rbuf, void* ws, void* stream) {
  MGP_CHECK_ARG(rbuf != nullptr, "cg_dist_norm2: rbuf is null");
  return cg_init<double>(b, ldb, (double*)b, (double*)b, (double*)b, ldb, n, ncols, 0., 0., 0., 1, 0, state, ws, (cudaStream_t)stream, 1, rbuf);
}
int mgp_cg_dist_init_f32(const float* b, int64_t ldb, float* x, float* r, float* p, int64_t ld, int64_t n, int32_t ncols,
                         float* state, float* rbuf, void* ws, void* stream) {
  MGP_CHECK_ARG(rbuf != nullptr, "cg_dist_init: rbuf is null");
  return cg_init<float>(b, ldb, x, r, p, ld, n, ncols, 0.f, 0.f, 0.f, 1, 0, state, ws, (cudaStream_t)stream, 2, rbuf);
}
int mgp_cg_dist_init_f64(const double* b, int64_t ldb, double* x, double* r, double* p, int64_t ld, int64_t n, int32_t ncols,
                         double* state, double* rbuf, void* ws, void* stream) {
  MGP_CHECK_ARG(rbuf != nullptr, "cg_dist_init: rbuf is null");
  return cg_init<double>(b, ldb, x, r, p, ld, n, ncols, 0., 0., 0., 1, 0, state, ws, (cudaStream_t)stream, 2, rbuf);
}
int mgp_cg_dist_update_f32(float* x, float* r, const float* p, const float* v, int64_t ld, int64_t n, int32_t ncols,
                           float* state, float* rbuf, void* ws, void* stream) {
  MGP_CHECK_ARG(rbuf != nullptr, "cg_dist_update: rbuf is null");
  return cg_update<float>(x, r, p, v, ld, n, ncols, state, nullptr, 0, ws, (cudaStream_t)stream, rbuf);
}
int mgp_cg_dist_update_f64(double* x, double* r, const double* p, const double* v, int64_t ld, int64_t n, int32_t ncols,
                           double* state, double* rbuf, void* ws, void* stream) {
  MGP_CHECK_ARG(rbuf != nullptr, "cg_dist_update: rbuf is null");
  return cg_update<double>(x, r, p, v, ld, n, ncols, state, nullptr, 0, ws, (cudaStream_t)stream, rbuf);
}
int mgp_cg_dist_scalars_f32(float* state, const float* rbuf, int32_t ncols, int32_t what, float tolerance, float eps,
                            float stop_updating_after, int32_t max_iter, int32_t n_tridiag_iter, float* hist, int32_t max_hist,
                            void* stream) {
  MGP_CHECK_ARG(state && rbuf && ncols > 0 && ncols <= kCgMaxCols && what >= 0 && what <= 2, "cg_dist_scalars: bad arguments");
  cg_scalars_kernel<float><<<1, kCgBlock, 0, (cudaStream_t)stream>>>(state, rbuf, ncols, what, tolerance, eps, stop_updating_after, max_iter, n_tridiag_iter, hist, max_hist);
  MGP_LAUNCH_CHECK();
  return MGP_OK;
}
int mgp_cg_dist_scalars_f64(double* state, const double* rbuf, int32_t ncols, int32_t what, double tolerance, double eps,
                            double stop_updating_after, int32_t max_iter, int32_t n_tridiag_iter, double* hist, int32_t max_hist,
                            void* stream) {
  MGP_CHECK_ARG(state && rbuf && ncols > 0 && ncols <= kCgMaxCols && what >= 0 && what <= 2, "cg_dist_scalars: bad arguments");
  cg_scalars_kernel<double><<<1, kCgBlock, 0, (cudaStream_t)stream>>>(state, rbuf, ncols, what, tolerance, eps, stop_updating_after, max_iter, n_tridiag_iter, hist, max_hist);
  MGP_LAUNCH_CHECK();
  return MGP_OK;
}
int mgp_cg_peer_scalars_f32(float* state, const float* rbuf, int32_t ncols, int32_t what, float tolerance, float eps,
                            float stop_updating_after, int32_t max_iter, int32_t n_tridiag_iter, float* hist, int32_t max_hist,
                            void* red_ptrs, void* flag_ptrs, void* epoch_ctr, int32_t rank, int32_t world, void* stream) {
  MGP_CHECK_ARG(state && rbuf && red_ptrs && flag_ptrs && epoch_ctr && ncols > 0 && ncols <= kCgMaxCols && what >= 0 && what <= 3 &&
                    world >= 1 && world <= 32 && rank >= 0 && rank < world, "cg_peer_scalars: bad arguments");
  cg_peer_scalars_kernel<float><<<1, kCgBlock, 0, (cudaStream_t)stream>>>(state, rbuf, ncols, what, tolerance, eps, stop_updating_after,
      max_iter, n_tridiag_iter, hist, max_hist, (float* const*)red_ptrs, (unsigned int* const*)flag_ptrs, (unsigned int*)epoch_ctr, rank, world);
  MGP_LAUNCH_CHECK();
  return MGP_OK;
}
int mgp_cg_peer_scalars_f64(double* state, const double* rbuf, int32_t ncols, int32_t what, double tolerance, double eps,
                            double stop_updating_after, int32_t max_iter, int32_t n_tridiag_iter, double* hist, int32_t max_hist,
                            void* red_ptrs, void* flag_ptrs, void* epoch_ctr, int32_t rank, int32_t world, void* stream) {
  MGP_CHECK_ARG(state && rbuf && red_ptrs && flag_ptrs && epoch_ctr && ncols > 0 && ncols <= kCgMaxCols && what >= 0 && what <= 3 &&
                    world >= 1 && world <= 32 && rank >= 0 && rank < world, "cg_peer_scalars: bad arguments");
  cg_peer_scalars_kernel<double><<<1, kCgBlock, 0, (cudaStream_t)stream>>>(state, rbuf, ncols, what, tolerance, eps, stop_updating_after,
      max_iter, n_tridiag_iter, hist, max_hist, (double* const*)red_ptrs, (unsigned int* const*)flag_ptrs, (unsigned int*)epoch_ctr, rank, world);
  MGP_LAUNCH_CHECK();
  return MGP_OK;
}
int mgp_cg_peer_rupdate_f32(float* r, const float* v, int64_t ld, int64_t n, int32_t ncols, float* state, const float* pap_local,
                            void* ws, void* red_ptrs, void* flag2_ptrs, void* flag3_ptrs, int32_t rank, int32_t world, void* stream) {
  MGP_CHECK_ARG(r && v && state && pap_local && ws && red_ptrs && flag2_ptrs && flag3_ptrs && n > 0 && ncols > 0 && world >= 1 &&
                    world <= 32 && rank >= 0 && rank < world, "cg_peer_rupdate: bad arguments");
  if (!cg_vec_ok<float>(ld, r, v, nullptr, nullptr)) return MGP_EUNSUPPORTED;
  int64_t g = ceil_div(n * ld / V16<float>::N, (int64_t)kCgBlock * 4); if (g > kNumSMs * 8) g = kNumSMs * 8; if (g < 1) g = 1;
  cg_peer_rupdate_kernel<float><<<(unsigned)g, kCgBlock, 0, (cudaStream_t)stream>>>(r, v, (int)ld, n, ncols, state, pap_local, ws,
      (float* const*)red_ptrs, (unsigned int* const*)flag2_ptrs, (unsigned int* const*)flag3_ptrs, rank, world);
  MGP_LAUNCH_CHECK();
  return MGP_OK;
}
int mgp_cg_peer_rupdate_f64(double* r, const double* v, int64_t ld, int64_t n, int32_t ncols, double* state, const double* pap_local,
                            void* ws, void* red_ptrs, void* flag2_ptrs, void* flag3_ptrs, int32_t rank, int32_t world, void* stream) {
  MGP_CHECK_ARG(r && v && state && pap_local && ws && red_ptrs && flag2_ptrs && flag3_ptrs && n > 0 && ncols > 0 && world >= 1 &&
                    world <= 32 && rank >= 0 && rank < world, "cg_peer_rupdate: bad arguments");
  if (!cg_vec_ok<double>(ld, r, v, nullptr, nullptr)) return MGP_EUNSUPPORTED;
  int64_t g = ceil_div(n * ld / V16<double>::N, (int64_t)kCgBlock * 4); if (g > kNumSMs * 8) g = kNumSMs * 8; if (g < 1) g = 1;
  cg_peer_rupdate_kernel<double><<<(unsigned)g, kCgBlock, 0, (cudaStream_t)stream>>>(r, v, (int)ld, n, ncols, state, pap_local, ws,
      (double* const*)red_ptrs, (unsigned int* const*)flag2_ptrs, (unsigned int* const*)flag3_ptrs, rank, world);
  MGP_LAUNCH_CHECK();
  return MGP_OK;
}
int mgp_cg_peer_pxupdate_f32(float* x, float* p, const float* r, int64_t ld, int64_t n, int32_t ncols, float* state, float* hist,
                             int32_t max_hist, void* ws, void* red_ptrs, void* flag3_ptrs, int32_t rank, int32_t world, void* stream) {
  MGP_CHECK_ARG(x && p && r && state && ws && red_ptrs && flag3_ptrs && n > 0 && ncols > 0 && world >= 1 && world <= 32 && rank >= 0 &&
                    rank < world, "cg_peer_pxupdate: bad arguments");
  if (!cg_vec_ok<float>(ld, x, p, r, nullptr)) return MGP_EUNSUPPORTED;
  int64_t g = ceil_div(n * ld / V16<float>::N, (int64_t)kCgBlock * 4); if (g > kNumSMs * 8) g = kNumSMs * 8; if (g < 1) g = 1;
  cg_peer_pxupdate_kernel<float><<<(unsigned)g, kCgBlock, 0, (cudaStream_t)stream>>>(x, p, r, (int)ld, n, ncols, state, hist, max_hist, ws,
      (float* const*)red_ptrs, (unsigned int* const*)flag3_ptrs, rank, world);
  MGP_LAUNCH_CHECK();
  return MGP_OK;
}
int mgp_cg_peer_pxupdate_f64(double* x, double* p, const double* r, int64_t ld, int64_t n, int32_t ncols, double* state, double* hist,
                             int32_t max_hist, void* ws, void* red_ptrs, void* flag3_ptrs, int32_t rank, int32_t world, void* stream) {
  MGP_CHECK_ARG(x && p && r && state && ws && red_ptrs && flag3_ptrs && n > 0 && ncols > 0 && world >= 1 && world <= 32 && rank >= 0 &&
                    rank < world, "cg_peer_pxupdate: bad arguments");
  if (!cg_vec_ok<double>(ld, x, p, r, nullptr)) return MGP_EUNSUPPORTED;
  int64_t g = ceil_div(n * ld / V16<double>::N, (int64_t)kCgBlock * 4); if (g > kNumSMs * 8) g = kNumSMs * 8; if (g < 1) g = 1;
  cg_peer_pxupdate_kernel<double><<<(unsigned)g, kCgBlock, 0, (cudaStream_t)stream>>>(x, p, r, (int)ld, n, ncols, state, hist, max_hist, ws,
      (double* const*)red_ptrs, (unsigned int* const*)flag3_ptrs, rank, world);
  MGP_LAUNCH_CHECK();
  return MGP_OK;
}
int mgp_cg_peer_cgstep_f32(float* x, float* r, float* p, float* s, const float* w, int64_t ld, int64_t n, int32_t ncols, float* state,
                           float* hist, int32_t max_hist, void* ws, float* gamma_loc, const float* delta_loc, void* red_ptrs, void* dflag_ptrs,
                           void* rflag_ptrs, int32_t rank, int32_t world, void* stream) {
  MGP_CHECK_ARG(x && r && p && s && w && state && ws && gamma_loc && red_ptrs && dflag_ptrs && n > 0 && ncols > 0 &&
                    world >= 1 && world <= 32 && rank >= 0 && rank < world, "cg_peer_cgstep: bad arguments");
  if (!cg_vec_ok<float>(ld, x, r, p, s) || (((uintptr_t)w) % 16) != 0) return MGP_EUNSUPPORTED;
  int64_t g = ceil_div(n * ld / V16<float>::N, (int64_t)kCgBlock * 4); if (g > kNumSMs * 8) g = kNumSMs * 8; if (g < 1) g = 1;
  MGP_CUDA(launch_pdl(cg_peer_cgstep_kernel<float>, dim3((unsigned)g), dim3(kCgBlock), 0, (cudaStream_t)stream, x, r, p, s, w, (int)ld, n, ncols,
                      state, hist, max_hist, ws, gamma_loc, delta_loc, (float* const*)red_ptrs, (unsigned int* const*)dflag_ptrs,
                      (unsigned int* const*)rflag_ptrs, rank, world));
  MGP_LAUNCH_CHECK();
  return MGP_OK;
}
int mgp_cg_peer_cgstep_f64(double* x, double* r, double* p, double* s, const double* w, int64_t ld, int64_t n, int32_t ncols, double* state,
                           double* hist, int32_t max_hist, void* ws, double* gamma_loc, const double* delta_loc, void* red_ptrs, void* dflag_ptrs,
                           void* rflag_ptrs, int32_t rank, int32_t world, void* stream) {
  MGP_CHECK_ARG(x && r && p && s && w && state && ws && gamma_loc && red_ptrs && dflag_ptrs && n > 0 && ncols > 0 &&
                    world >= 1 && world <= 32 && rank >= 0 && rank < world, "cg_peer_cgstep: bad arguments");
  if (!cg_vec_ok<double>(ld, x, r, p, s) || (((uintptr_t)w) % 16) != 0) return MGP_EUNSUPPORTED;
  int64_t g = ceil_div(n * ld / V16<double>::N, (int64_t)kCgBlock * 4); if (g > kNumSMs * 8) g = kNumSMs * 8; if (g < 1) g = 1;
  MGP_CUDA(launch_pdl(cg_peer_cgstep_kernel<double>, dim3((unsigned)g), dim3(kCgBlock), 0, (cudaStream_t)stream, x, r, p, s, w, (int)ld, n, ncols,
                      state, hist, max_hist, ws, gamma_loc, delta_loc, (double* const*)red_ptrs, (unsigned int* const*)dflag_ptrs,
                      (unsigned int* const*)rflag_ptrs, rank, world));
  MGP_LAUNCH_CHECK();
  return MGP_OK;
}
int mgp_peer_publish(void* flag_ptrs, uint32_t value, int32_t rank, int32_t world, void* stream) {
  MGP_CHECK_ARG(flag_ptrs && world >= 1 && world <= 32 && rank >= 0 && rank < world, "peer_publish: bad arguments");
  peer_publish_kernel<<<1, 32, 0, (cudaStream_t)stream>>>((unsigned int* const*)flag_ptrs, value, rank, world);
  MGP_LAUNCH_CHECK();
  return MGP_OK;
}
int mgp_peer_barrier(void* flag_ptrs, void* epoch_ctr, int32_t rank, int32_t world, void* stream) {
  MGP_CHECK_ARG(flag_ptrs && epoch_ctr && world >= 1 && world <= 32 && rank >= 0 && rank < world, "peer_barrier: bad arguments");
  peer_barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>((unsigned int* const*)flag_ptrs, (unsigned int*)epoch_ctr, rank, world);
  MGP_LAUNCH_CHECK();
  return MGP_OK;
}
int mgp_cg_pupdate_f32(float* p, const float* r, int64_t ld, int64_t n, int32_t ncols, const float* state, void* stream) {
  MGP_CHECK_ARG(p && r && state && n > 0 && ncols > 0 && ld >= ncols, "cg_pupdate: bad arguments");
  const int64_t total = n * ld;
  if (cg_vec_ok<float>(ld, p, r, nullptr, nullptr)) {
    int64_t gv = ceil_div(total / V16<float>::N, (int64_t)kCgBlock * 4); if (gv > kNumSMs * 8) gv = kNumSMs * 8; if (gv < 1) gv = 1;
    cg_pupdate_vec_kernel<float><<<(unsigned)gv, kCgBlock, 0, (cudaStream_t)stream>>>(p, r, (int)ld, n, ncols, state);
    MGP_LAUNCH_CHECK();
    return MGP_OK;
  }
  int64_t g = ceil_div(total, kCgBlock); if (g > kNumSMs * 16) g = kNumSMs * 16;
  cg_pupdate_kernel<float><<<(unsigned)g, kCgBlock, 0, (cudaStream_t)stream>>>(p, r, ld, n, ncols, state);
  MGP_LAUNCH_CHECK();
  return MGP_OK;
}
int mgp_cg_pupdate_f64(double* p, const double* r, int64_t ld, int64_t n, int32_t ncols, const double* state, void* stream) {
  MGP_CHECK_ARG(p && r && state && n > 0 && ncols > 0 && ld >= ncols, "cg_pupdate: bad arguments");
  const int64_t total = n * ld;
  if (cg_vec_ok<double>(ld, p, r, nullptr, nullptr)) {
    int64_t gv = ceil_div(total / V16<double>::N, (int64_t)kCgBlock * 4); if (gv > kNumSMs * 8) gv = kNumSMs * 8; if (gv < 1) gv = 1;
    cg_pupdate_vec_kernel<double><<<(unsigned)gv, kCgBlock, 0, (cudaStream_t)stream>>>(p, r, (int)ld, n, ncols, state);
    MGP_LAUNCH_CHECK();
    return MGP_OK;
  }
  int64_t g = ceil_div(total, kCgBlock); if (g > kNumSMs * 16) g = kNumSMs * 16;
  cg_pupdate_kernel<double><<<(unsigned)g, kCgBlock, 0, (cudaStream_t)stream>>>(p, r, ld, n, ncols, state);
  MGP_LAUNCH_CHECK();
  return MGP_OK;
}
int mgp_cg_finalize_f32(const float* x, int64_t ld, float* out, int64_t ldo, int64_t n, int32_t ncols, const float* state,
                        void* stream) {
  MGP_CHECK_ARG(x && out && state && n > 0 && ncols > 0 && ld >= ncols && ldo >= ncols, "cg_finalize: bad arguments");
  int64_t g = ceil_div(n * ncols, kCgBlock); if (g > kNumSMs * 16) g = kNumSMs * 16;
  cg_finalize_kernel<float><<<(unsigned)g, kCgBlock, 0, (cudaStream_t)stream>>>(x, ld, out, ldo, n, ncols, state);
  MGP_LAUNCH_CHECK();
  return MGP_OK;
}
int mgp_cg_finalize_f64(const double* x, int64_t ld, double* out, int64_t ldo, int64_t n, int32_t ncols,
                        const double* state, void* stream) {
  MGP_CHECK_ARG(x && out && state && n > 0 && ncols > 0 && ld >= ncols && ldo >= ncols, "cg_finalize: bad arguments");
  int64_t g = ceil_div(n * ncols, kCgBlock); if (g > kNumSMs * 16) g = kNumSMs * 16;
  cg_finalize_kernel<double><<<(unsigned)g, kCgBlock, 0, (cudaStream_t)stream>>>(x, ld, out, ldo, n, ncols, state);
  MGP_LAUNCH_CHECK();
  return MGP_OK;
}

}  // extern "C"
