// Fused graph-Laplacian / Matern-step SpMM over the row-major directed structure, and its SDDMM backward.
//
//   Y = post .* ( (diag + shift) .* (pre .* X)  -  A (pre .* X) )
//
// Reference: GraphLaplacianOperator._matmul, manifold_gp/operators/graph_laplacian_operator.py:108-124
// (out = vec*diag; out -= spmm(idx, triu, vec); out -= spmm(idx.flip, triu, vec); D^{+-1/2} scalings for the
// random-walk normalisation) and one step of PrecisionMaternOperator._matmul, precision_matern_operator.py:28-32.
// The reference materialises two [M, C] temporaries per matvec and scatters with atomics; here each output row is
// one deterministic sub-warp reduction over its CSR row (both directions of every edge are stored), the diagonal,
// the Matern shift and the degree scalings are fused, and an optional dot-product epilogue (CG's p^T A p,
// Lanczos' alpha) removes a further pass over the vectors.
//
// Kernel shape (v1, "csr-subwarp"): LPR lanes cooperate on one row; each nonzero is handled by LPN adjacent lanes
// that own VEC consecutive right-hand-side columns each (128-bit gathers of X rows when VEC > 1).  Persistent
// grid-stride over row blocks so the dot epilogue needs only gridDim.x partials.
#include "common.cuh"
#include "spmm_common.cuh"

namespace mgp {

template <typename T>
struct SpmmArgs {
  const int* rowptr;
  const int* col;
  const T* a;
  const T* diag;
  const T* shift;  // device scalar or null
  const T* pre;    // [n] or null
  const T* post;   // [n] or null
  const T* x;
  int64_t ldx;
  T* y;
  int64_t ldy;
  int64_t n;
  int c0;  // first column of this pass
  int cw;  // valid columns in this pass
  const T* dot_with;  // ld = ldx, or null
  T* dot_out;         // [ncols] (already offset by nothing: indexed c0 + c)
  T* partials;        // [gridDim.x, cw]
  unsigned int* counter;
};

constexpr int kSpmmBlock = 256;

template <typename T, int VEC, int LPN, int LPR>
__global__ void __launch_bounds__(kSpmmBlock)
lap_spmm_csr_kernel(const SpmmArgs<T> g) {
  static_assert(LPR % LPN == 0 && 32 % LPR == 0, "bad lane mapping");
  constexpr int SPR = LPR / LPN;               // nonzeros of one row in flight per iteration
  constexpr int ROWS_PER_BLOCK = kSpmmBlock / LPR;
  constexpr int CWMAX = LPN * VEC;
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int l = lane % LPR;     // lane within the row group
  const int slot = l / LPN;     // which nonzero of the group
  const int cl = l % LPN;       // which column group
  const int cbase = g.c0 + cl * VEC;
  const bool col_ok = (VEC > 1) ? true : (cl < g.cw);
  const T shift = g.shift ? *g.shift : T(0);

  T dsum[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) dsum[v] = T(0);

  for (int64_t row_base = (int64_t)blockIdx.x * ROWS_PER_BLOCK; row_base < g.n;
       row_base += (int64_t)gridDim.x * ROWS_PER_BLOCK) {
    const int64_t row = row_base + tid / LPR;
    int p0 = 0, p1 = 0;
    if (row < g.n) {
      p0 = __ldg(g.rowptr + row);
      p1 = __ldg(g.rowptr + row + 1);
    }
    T acc[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) acc[v] = T(0);
    if (col_ok) {
#pragma unroll 4
      for (int p = p0 + slot; p < p1; p += SPR) {
        const int j = __ldg(g.col + p);
        T w = __ldg(g.a + p);
        if (g.pre) w *= __ldg(g.pre + j);
        const Vec<T, VEC> xv = ldg_vec<T, VEC>(g.x + (int64_t)j * g.ldx + cbase);
#pragma unroll
        for (int v = 0; v < VEC; ++v) acc[v] = fma(w, xv.v[v], acc[v]);
      }
    }
    // reduce the SPR partial sums of the row (lanes that differ in the slot bits)
#pragma unroll
    for (int v = 0; v < VEC; ++v) acc[v] = subwarp_sum(acc[v], LPN, LPR);
    if (slot == 0 && row < g.n && col_ok) {
      Vec<T, VEC> xi = ldg_vec<T, VEC>(g.x + row * g.ldx + cbase);
      T d = __ldg(g.diag + row) + shift;
      if (g.pre) d *= __ldg(g.pre + row);
      const T po = g.post ? __ldg(g.post + row) : T(1);
      Vec<T, VEC> out;
#pragma unroll
      for (int v = 0; v < VEC; ++v) out.v[v] = po * (d * xi.v[v] - acc[v]);
      st_vec<T, VEC>(g.y + row * g.ldy + cbase, out);
      if (g.dot_out) {
        const Vec<T, VEC> dw = ldg_vec<T, VEC>(g.dot_with + row * g.ldx + cbase);
#pragma unroll
        for (int v = 0; v < VEC; ++v) dsum[v] = fma(dw.v[v], out.v[v], dsum[v]);
      }
    }
  }

  if (g.dot_out) {
    // block partial per column: lanes with equal cl across the warp, then across warps (fixed order)
    __shared__ T sm[kSpmmBlock / 32][CWMAX];
#pragma unroll
    for (int v = 0; v < VEC; ++v) dsum[v] = subwarp_sum(dsum[v], LPN, 32);
    if (lane < LPN) {
#pragma unroll
      for (int v = 0; v < VEC; ++v) sm[tid >> 5][lane * VEC + v] = dsum[v];
    }
    __syncthreads();
    if (tid < g.cw) {
      T s = T(0);
#pragma unroll
      for (int w = 0; w < kSpmmBlock / 32; ++w) s += sm[w][tid];
      g.partials[(int64_t)blockIdx.x * g.cw + tid] = s;
    }
    if (last_block_ticket(g.counter)) {
      // fixed-order reduction over blocks: thread t handles column t % cw, strided over blocks, then tree in smem
      __shared__ T red[kSpmmBlock];
      const int c = tid % g.cw;
      const int lanes_per_col = kSpmmBlock / g.cw;  // cw <= 32
      const int r = tid / g.cw;
      T s = T(0);
      if (r < lanes_per_col)
        for (int b = r; b < (int)gridDim.x; b += lanes_per_col) s += __ldcg(g.partials + (int64_t)b * g.cw + c);
      red[tid] = (r < lanes_per_col) ? s : T(0);
      __syncthreads();
      if (tid < g.cw) {
        T t = T(0);
        for (int rr = 0; rr < lanes_per_col; ++rr) t += red[rr * g.cw + tid];
        g.dot_out[g.c0 + tid] = t;
      }
    }
  }
}

// ---- SDDMM -----------------------------------------------------------------------------------------------------
// One warp per row; LPN lanes per nonzero, each lane striding over the columns.
template <typename T, int LPN>
__global__ void __launch_bounds__(256)
lap_sddmm_kernel(const int* __restrict__ rowptr, const int* __restrict__ col, const T* __restrict__ pre,
                 const T* __restrict__ post, const T* __restrict__ gy, int64_t ldgy, const T* __restrict__ x,
                 int64_t ldx, int64_t n, int ncols, T* __restrict__ g_a, T* __restrict__ g_diag) {
  constexpr int SLOTS = 32 / LPN;
  const int lane = threadIdx.x & 31;
  const int slot = lane / LPN;
  const int cl = lane % LPN;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t row = warp; row < n; row += nwarps) {
    const int p0 = __ldg(rowptr + row), p1 = __ldg(rowptr + row + 1);
    const T po = post ? __ldg(post + row) : T(1);
    const T* gyr = gy + row * ldgy;
    for (int pb = p0; pb < p1; pb += SLOTS) {   // warp-uniform trip count
      const int p = pb + slot;
      T s = T(0);
      if (p < p1) {
        const int j = __ldg(col + p);
        const T pj = pre ? __ldg(pre + j) : T(1);
        const T* xr = x + (int64_t)j * ldx;
        for (int c = cl; c < ncols; c += LPN) s = fma(__ldg(gyr + c), __ldg(xr + c), s);
        s *= -(po * pj);
      }
      s = subwarp_sum(s, 1, LPN);
      if (p < p1 && cl == 0) g_a[p] = s;
    }
    // diagonal
    T d = T(0);
    for (int c = lane; c < ncols; c += 32) d = fma(__ldg(gyr + c), __ldg(x + row * ldx + c), d);
    d = warp_sum(d);
    if (lane == 0) g_diag[row] = d * po * (pre ? __ldg(pre + row) : T(1));
  }
}

// 128-bit form for column counts that are whole 16-byte chunks (ncols = VEC * LPN, LPN a power of two <= 8): LPN lanes per
// nonzero, each holding its chunk of the row's gy in registers for the whole row and loading one 16-byte chunk of x[col] per
// entry -- 32 / LPN entries per warp iteration instead of 2 (measured at cfg-C, C = 16: the scalar kernel above took 1.93 ms =
// 0.03 of the HBM peak on its algorithmic bytes, bound by the latency of its dependent 4-byte gathers).
template <typename T, int VEC, int LPN>
__global__ void __launch_bounds__(256)
lap_sddmm_vec_kernel(const int* __restrict__ rowptr, const int* __restrict__ col, const T* __restrict__ pre,
                     const T* __restrict__ post, const T* __restrict__ gy, int64_t ldgy, const T* __restrict__ x,
                     int64_t ldx, int64_t n, T* __restrict__ g_a, T* __restrict__ g_diag) {
  constexpr int SLOTS = 32 / LPN;
  const int lane = threadIdx.x & 31;
  const int slot = lane / LPN;
  const int cl = lane % LPN;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t row = warp; row < n; row += nwarps) {
    const int p0 = __ldg(rowptr + row), p1 = __ldg(rowptr + row + 1);
    const T po = post ? __ldg(post + row) : T(1);
    const Vec<T, VEC> g = ldg_vec<T, VEC>(gy + row * ldgy + cl * VEC);
    const Vec<T, VEC> xd = ldg_vec<T, VEC>(x + row * ldx + cl * VEC);
    for (int pb = p0; pb < p1; pb += SLOTS) {   // warp-uniform trip count
      const int p = pb + slot;
      T s = T(0);
      if (p < p1) {
        const int j = ld_stream(col + p);
        const T pj = pre ? __ldg(pre + j) : T(1);
        const Vec<T, VEC> xv = ldg_vec<T, VEC>(x + (int64_t)j * ldx + cl * VEC);
#pragma unroll
        for (int v = 0; v < VEC; ++v) s = fma(g.v[v], xv.v[v], s);
        s *= -(po * pj);
      }
      s = subwarp_sum(s, 1, LPN);
      if (p < p1 && cl == 0) g_a[p] = s;
    }
    T d = T(0);
#pragma unroll
    for (int v = 0; v < VEC; ++v) d = fma(g.v[v], xd.v[v], d);
    d = subwarp_sum(d, 1, LPN);                  // every slot holds the same chunks: the sum over one slot's lanes is the row's
    if (lane == 0) g_diag[row] = d * po * (pre ? __ldg(pre + row) : T(1));
  }
}

template <typename T, int VEC, int LPN, int LPR>
static int launch_spmm(const SpmmArgs<T>& g, cudaStream_t st) {
  constexpr int ROWS_PER_BLOCK = kSpmmBlock / LPR;
  int64_t blocks = ceil_div(g.n, ROWS_PER_BLOCK);
  const int64_t cap = (int64_t)kNumSMs * 8;   // persistent: 8 CTAs of 256 threads per SM
  if (blocks > cap) blocks = cap;
  lap_spmm_csr_kernel<T, VEC, LPN, LPR><<<(unsigned)blocks, kSpmmBlock, 0, st>>>(g);
  MGP_LAUNCH_CHECK();
  return MGP_OK;
}

struct DotWs {
  unsigned int counter[64];  // first word used; padded
};

template <typename T>
static int lap_spmm(const int* rowptr, const int* col, const T* a, const T* diag, const T* shift, const T* pre,
                    const T* post, const T* x, int64_t ldx, T* y, int64_t ldy, int64_t n, int ncols, const T* dot_with,
                    T* dot_out, void* dot_ws, cudaStream_t st) {
  MGP_CHECK_ARG(rowptr && col && a && diag && x && y, "lap_spmm: null pointer");
  MGP_CHECK_ARG(n > 0 && ncols > 0 && ldx >= ncols && ldy >= ncols, "lap_spmm: bad shape n=%lld ncols=%d ldx=%lld ldy=%lld",
                (long long)n, ncols, (long long)ldx, (long long)ldy);
  MGP_CHECK_ARG(x != y, "lap_spmm: X and Y must not alias");
  MGP_CHECK_ARG((dot_out == nullptr) || (dot_with && dot_ws), "lap_spmm: dot epilogue needs dot_with and dot_ws");
  constexpr int VECW = sizeof(T) == 4 ? 4 : 2;
  const bool aligned = (ldx % VECW == 0) && (ldy % VECW == 0) && (((uintptr_t)x) % 16 == 0) && (((uintptr_t)y) % 16 == 0) &&
                       (dot_with == nullptr || ((uintptr_t)dot_with) % 16 == 0);
  SpmmArgs<T> g;
  g.rowptr = rowptr; g.col = col; g.a = a; g.diag = diag; g.shift = shift; g.pre = pre; g.post = post;
  g.x = x; g.ldx = ldx; g.y = y; g.ldy = ldy; g.n = n; g.dot_with = dot_out ? dot_with : nullptr; g.dot_out = dot_out;
  g.counter = dot_out ? reinterpret_cast<unsigned int*>(dot_ws) : nullptr;
  g.partials = dot_out ? reinterpret_cast<T*>(reinterpret_cast<char*>(dot_ws) + 256) : nullptr;
  int c0 = 0;
  while (c0 < ncols) {
    const int rem = ncols - c0;
    int rc;
    g.c0 = c0;
    if (aligned && c0 % VECW == 0 && rem >= VECW) {
      if constexpr (sizeof(T) == 4) {
        if (rem >= 16) { g.cw = 16; rc = launch_spmm<T, 4, 4, 16>(g, st); }
        else if (rem >= 8) { g.cw = 8; rc = launch_spmm<T, 4, 2, 8>(g, st); }
        else { g.cw = 4; rc = launch_spmm<T, 4, 1, 8>(g, st); }
      } else {
        if (rem >= 16) { g.cw = 16; rc = launch_spmm<T, 2, 8, 32>(g, st); }
        else if (rem >= 8) { g.cw = 8; rc = launch_spmm<T, 2, 4, 16>(g, st); }
        else if (rem >= 4) { g.cw = 4; rc = launch_spmm<T, 2, 2, 8>(g, st); }
        else { g.cw = 2; rc = launch_spmm<T, 2, 1, 8>(g, st); }
      }
    } else {
      // scalar path: one column per lane, up to 32 columns per pass
      const int cw = rem > 32 ? 32 : rem;
      g.cw = cw;
      if (cw == 1) rc = launch_spmm<T, 1, 1, 8>(g, st);
      else if (cw == 2) rc = launch_spmm<T, 1, 2, 8>(g, st);
      else if (cw <= 4) rc = launch_spmm<T, 1, 4, 16>(g, st);
      else if (cw <= 8) rc = launch_spmm<T, 1, 8, 32>(g, st);
      else if (cw <= 16) rc = launch_spmm<T, 1, 16, 32>(g, st);
      else rc = launch_spmm<T, 1, 32, 32>(g, st);
    }
    if (rc != MGP_OK) return rc;
    c0 += g.cw;
  }
  return MGP_OK;
}

template <typename T>
static int lap_sddmm(const int* rowptr, const int* col, const T* pre, const T* post, const T* gy, int64_t ldgy, const T* x,
                     int64_t ldx, int64_t n, int ncols, T* g_a, T* g_diag, cudaStream_t st) {
  MGP_CHECK_ARG(rowptr && col && gy && x && g_a && g_diag, "lap_sddmm: null pointer");
  MGP_CHECK_ARG(n > 0 && ncols > 0 && ldx >= ncols && ldgy >= ncols, "lap_sddmm: bad shape");
  int64_t blocks = ceil_div(n, 256 / 32);
  const int64_t cap = (int64_t)kNumSMs * 8;
  if (blocks > cap) blocks = cap;
  constexpr int VEC = 16 / (int)sizeof(T);
  const bool vec_ok = ncols % VEC == 0 && ldgy % VEC == 0 && ldx % VEC == 0 && ((uintptr_t)gy) % 16 == 0 && ((uintptr_t)x) % 16 == 0;
  const int lpn = vec_ok ? ncols / VEC : 0;
  if (lpn == 1 || lpn == 2 || lpn == 4 || lpn == 8) {
    if (lpn == 1) lap_sddmm_vec_kernel<T, VEC, 1><<<(unsigned)blocks, 256, 0, st>>>(rowptr, col, pre, post, gy, ldgy, x, ldx, n, g_a, g_diag);
    else if (lpn == 2) lap_sddmm_vec_kernel<T, VEC, 2><<<(unsigned)blocks, 256, 0, st>>>(rowptr, col, pre, post, gy, ldgy, x, ldx, n, g_a, g_diag);
    else if (lpn == 4) lap_sddmm_vec_kernel<T, VEC, 4><<<(unsigned)blocks, 256, 0, st>>>(rowptr, col, pre, post, gy, ldgy, x, ldx, n, g_a, g_diag);
    else lap_sddmm_vec_kernel<T, VEC, 8><<<(unsigned)blocks, 256, 0, st>>>(rowptr, col, pre, post, gy, ldgy, x, ldx, n, g_a, g_diag);
    MGP_LAUNCH_CHECK();
    return MGP_OK;
  }
  if (ncols == 1) lap_sddmm_kernel<T, 1><<<(unsigned)blocks, 256, 0, st>>>(rowptr, col, pre, post, gy, ldgy, x, ldx, n, ncols, g_a, g_diag);
  else if (ncols == 2) lap_sddmm_kernel<T, 2><<<(unsigned)blocks, 256, 0, st>>>(rowptr, col, pre, post, gy, ldgy, x, ldx, n, ncols, g_a, g_diag);
  else if (ncols <= 4) lap_sddmm_kernel<T, 4><<<(unsigned)blocks, 256, 0, st>>>(rowptr, col, pre, post, gy, ldgy, x, ldx, n, ncols, g_a, g_diag);
  else if (ncols <= 8) lap_sddmm_kernel<T, 8><<<(unsigned)blocks, 256, 0, st>>>(rowptr, col, pre, post, gy, ldgy, x, ldx, n, ncols, g_a, g_diag);
  else lap_sddmm_kernel<T, 16><<<(unsigned)blocks, 256, 0, st>>>(rowptr, col, pre, post, gy, ldgy, x, ldx, n, ncols, g_a, g_diag);
  MGP_LAUNCH_CHECK();
  return MGP_OK;
}

}  // namespace mgp

extern "C" {

size_t mgp_lap_spmm_dot_ws_bytes(int64_t n, int32_t ncols) {
  (void)n;
  // counter block + per-block partials of the widest pass (32 columns), 8 bytes per value
  return 256 + (size_t)mgp::kNumSMs * 8 * 32 * 8 + (size_t)ncols * 0;
}

int mgp_lap_spmm_f32(const int32_t* rowptr, const int32_t* col, const float* a, const float* diag, const float* shift,
                     const float* pre, const float* post, const float* x, int64_t ldx, float* y, int64_t ldy, int64_t n,
                     int32_t ncols, const float* dot_with, float* dot_out, void* dot_ws, void* stream) {
  return mgp::lap_spmm<float>(rowptr, col, a, diag, shift, pre, post, x, ldx, y, ldy, n, ncols, dot_with, dot_out, dot_ws,
                              (cudaStream_t)stream);
}

int mgp_lap_spmm_f64(const int32_t* rowptr, const int32_t* col, const double* a, const double* diag, const double* shift,
                     const double* pre, const double* post, const double* x, int64_t ldx, double* y, int64_t ldy,
                     int64_t n, int32_t ncols, const double* dot_with, double* dot_out, void* dot_ws, void* stream) {
  return mgp::lap_spmm<double>(rowptr, col, a, diag, shift, pre, post, x, ldx, y, ldy, n, ncols, dot_with, dot_out, dot_ws,
                               (cudaStream_t)stream);
}

int mgp_lap_sddmm_f32(const int32_t* rowptr, const int32_t* col, const float* pre, const float* post, const float* gy,
                      int64_t ldgy, const float* x, int64_t ldx, int64_t n, int32_t ncols, float* g_a, float* g_diag,
                      void* stream) {
  return mgp::lap_sddmm<float>(rowptr, col, pre, post, gy, ldgy, x, ldx, n, ncols, g_a, g_diag, (cudaStream_t)stream);
}

int mgp_lap_sddmm_f64(const int32_t* rowptr, const int32_t* col, const double* pre, const double* post, const double* gy,
                      int64_t ldgy, const double* x, int64_t ldx, int64_t n, int32_t ncols, double* g_a, double* g_diag,
                      void* stream) {
  return mgp::lap_sddmm<double>(rowptr, col, pre, post, gy, ldgy, x, ldx, n, ncols, g_a, g_diag, (cudaStream_t)stream);
}

}  // extern "C"
