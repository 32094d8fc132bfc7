// v2 SpMM: "tile-compacted" rows with the right-hand-side tile staged in shared memory.
//
//   Y = post .* ( (diag + shift) .* (pre .* X)  -  A (pre .* X) )          (same contract as lap_spmm.cu)
//
// Why: with 16 right-hand sides every nonzero gathers a 64-byte row of X.  The v1 kernel sends those gathers through
// L1 (one wavefront per distinct line): ncu shows l1tex at ~87 % and DRAM at ~20 % -- the kernel is bound by the
// load/store pipe, not by HBM.  After the space-filling-curve reordering a tile of R consecutive rows only touches
// its own R rows of X plus a small halo (2-D manifold: ~1-2x R rows), so the tile's slice of X fits in shared memory:
//   phase A  gather the R + H distinct X rows of the tile ONCE (coalesced 128-bit loads, `pre` folded in) and stage the
//            tile's 16-bit local column indices and values with 128-bit streaming loads (2 + w bytes per nonzero
//            instead of 4 + w: the index stream shrinks by half);
//   phase B  sub-warps walk their rows reading (index, value) as shared-memory broadcasts and X rows as conflict-light
//            128-bit shared loads; one deterministic shuffle reduction per row;
//   phase C  fused diagonal / Matern shift / `post`, 128-bit stores of Y, optional dot-product epilogue.
// Several CTAs are resident per SM so one tile's phase A overlaps another's phase B.
//
// Structure (built once per graph by the host, see graph.py::TileStructure): rowptr (CSR of the reordered graph),
// lcol[nnz] uint16 (local index: own rows 0..R-1, halo rows R..R+H-1), halo_ptr[T+1], halo_col[sum H].
// `xmap` / `ymap` (optional) translate the structure's row ids to the caller's row order for X / Y.
#include <stdlib.h>

#include "common.cuh"
#include "spmm_common.cuh"

namespace mgp {

constexpr int kTiledBlock = 256;

template <typename T>
struct TiledArgs {
  const int* rowptr;
  const unsigned short* lcol;
  const T* a;
  const T* diag;
  const int* halo_ptr;
  const int* halo_col;
  const T* shift;
  const T* pre;
  const T* post;
  const int* xmap;
  const int* ymap;
  const T* x;
  int64_t ldx;
  T* y;
  int64_t ldy;
  int64_t n;
  int ntiles;
  int lmax;   // max over tiles of R + H
  int nzcap;  // staging capacity in entries (multiple of 8, >= max tile nnz + 8)
  int c0, cw;
  const T* dot_with;
  T* dot_out;
  T* partials;
  unsigned int* counter;
  int dot_is_x;  // dot_with == x and no pre scaling: the row is already in shared memory
  int debug;     // development only: 1 = skip phase B, 3 = phase A only once per CTA
};

template <typename T, int VEC, int LPN, int LPR, int R>
__global__ void __launch_bounds__(kTiledBlock)
lap_spmm_tiled_kernel(const TiledArgs<T> g) {
  static_assert(LPR % LPN == 0 && 32 % LPR == 0, "bad lane mapping");
  constexpr int CW = LPN * VEC;                 // columns per pass == shared-memory row stride of the X tile
  constexpr int SPR = LPR / LPN;
  constexpr int ROWS_PER_ROUND = kTiledBlock / LPR;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* xs = reinterpret_cast<T*>(smem_raw);                                   // [lmax][CW]
  T* vs = xs + (size_t)g.lmax * CW;                                         // [nzcap]
  unsigned short* cs = reinterpret_cast<unsigned short*>(vs + g.nzcap);     // [nzcap]
  int* rp = reinterpret_cast<int*>(cs + g.nzcap);                           // [R + 4]
  T* dgs = reinterpret_cast<T*>(rp + R + 4);                                // [R]  diag + shift of the tile's rows
  T* pss = dgs + R;                                                         // [R]  post scaling of the tile's rows

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int l = lane % LPR;
  const int slot = l / LPN;
  const int cl = l % LPN;
  const int cbase = g.c0 + cl * VEC;
  const bool col_ok = (VEC > 1) ? true : (cl < g.cw);
  const T shift = g.shift ? *g.shift : T(0);

  T dsum[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) dsum[v] = T(0);

  for (int tile = blockIdx.x; tile < g.ntiles; tile += gridDim.x) {
    const int64_t row0 = (int64_t)tile * R;
    const int nrows = (int)min((int64_t)R, g.n - row0);
    const int p0 = __ldg(g.rowptr + row0);
    const int p1 = __ldg(g.rowptr + row0 + nrows);
    const int base = p0 & ~7;                      // 16-byte aligned start of the staged index / value streams
    const int cnt = p1 - base;
    const int h0 = __ldg(g.halo_ptr + tile);
    const int nh = __ldg(g.halo_ptr + tile + 1) - h0;

    // ---- phase A: stage indices + values (128-bit streaming loads), row offsets, and the X tile ------------------
    if (g.debug != 3 || tile == (int)blockIdx.x) {   // debug 3: phase A only for the first tile (stale but valid indices)
    for (int i = tid * 8; i < cnt; i += kTiledBlock * 8) {
      const int4 c8 = ld_stream_v4(reinterpret_cast<const int4*>(g.lcol + base + i));
      *reinterpret_cast<int4*>(cs + i) = c8;
      if constexpr (sizeof(T) == 4) {
        const int4 v0 = ld_stream_v4(reinterpret_cast<const int4*>(g.a + base + i));
        const int4 v1 = ld_stream_v4(reinterpret_cast<const int4*>(g.a + base + i + 4));
        *reinterpret_cast<int4*>(vs + i) = v0;
        *reinterpret_cast<int4*>(vs + i + 4) = v1;
      } else {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int4 v0 = ld_stream_v4(reinterpret_cast<const int4*>(g.a + base + i + 2 * u));
          *reinterpret_cast<int4*>(vs + i + 2 * u) = v0;
        }
      }
    }
    for (int i = tid; i <= nrows; i += kTiledBlock) rp[i] = __ldg(g.rowptr + row0 + i) - base;
    for (int i = tid; i < nrows; i += kTiledBlock) {
      dgs[i] = __ldg(g.diag + row0 + i) + shift;
      pss[i] = g.post ? __ldg(g.post + row0 + i) : T(1);
    }
    // X tile: ids first, then all gathers of the batch in flight together, then the shared-memory stores
    {
      constexpr int UNR = 4;
      const int nitems = (R + nh) * LPN;
      for (int it0 = tid; it0 < nitems; it0 += kTiledBlock * UNR) {
        int id[UNR], lr[UNR], cc[UNR];
        bool ok[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
          const int it = it0 + u * kTiledBlock;
          ok[u] = it < nitems;
          lr[u] = it / LPN;
          cc[u] = it % LPN;
          id[u] = 0;
          if (ok[u]) {
            if (lr[u] < R) {
              ok[u] = lr[u] < nrows;
              id[u] = (int)row0 + lr[u];
            } else {
              id[u] = __ldg(g.halo_col + h0 + lr[u] - R);
            }
            if (VEC == 1 && cc[u] >= g.cw) ok[u] = false;
          }
        }
        int64_t src[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) src[u] = (ok[u] && g.xmap) ? (int64_t)__ldg(g.xmap + id[u]) : (int64_t)id[u];
        Vec<T, VEC> v[UNR];
        T pj[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
          if (ok[u]) {
            v[u] = ldg_vec<T, VEC>(g.x + src[u] * g.ldx + g.c0 + cc[u] * VEC);
            pj[u] = g.pre ? __ldg(g.pre + id[u]) : T(1);
          }
        }
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
          if (ok[u]) {
            if (g.pre) {
#pragma unroll
              for (int e = 0; e < VEC; ++e) v[u].v[e] *= pj[u];
            }
            if constexpr (VEC == 1) xs[lr[u] * CW + cc[u]] = v[u].v[0];
            else *reinterpret_cast<Vec<T, VEC>*>(xs + lr[u] * CW + cc[u] * VEC) = v[u];
          }
        }
      }
    }
    }
    __syncthreads();

    // ---- phase B + C ---------------------------------------------------------------------------------------------
    for (int rb = 0; rb < (g.debug == 1 ? 0 : R); rb += ROWS_PER_ROUND) {       // block-uniform trip count (shuffles below need full warps)
      const int r = rb + tid / LPR;
      const bool writer = (slot == 0) && (r < nrows) && col_ok;
      int q0 = 0, q1 = 0;
      if (r < nrows) { q0 = rp[r]; q1 = rp[r + 1]; }
      const int64_t row = row0 + r;
      Vec<T, VEC> dw;
      if (g.dot_out && !g.dot_is_x && writer) {              // issue early: the latency hides behind the row loop
        const int64_t drow = g.xmap ? (int64_t)__ldg(g.xmap + row) : row;
        dw = ldg_vec<T, VEC>(g.dot_with + drow * g.ldx + cbase);
      }
      int64_t yrow = row;
      if (g.ymap && writer) yrow = (int64_t)__ldg(g.ymap + row);
      T acc[VEC];
#pragma unroll
      for (int v = 0; v < VEC; ++v) acc[v] = T(0);
      if (col_ok) {
#pragma unroll 4
        for (int p = q0 + slot; p < q1; p += SPR) {
          const int j = cs[p];
          const T w = vs[p];
          if constexpr (VEC == 1) {
            acc[0] = fma(w, xs[j * CW + cl], acc[0]);
          } else {
            const Vec<T, VEC> xv = *reinterpret_cast<const Vec<T, VEC>*>(xs + j * CW + cl * VEC);
#pragma unroll
            for (int v = 0; v < VEC; ++v) acc[v] = fma(w, xv.v[v], acc[v]);
          }
        }
      }
      if constexpr (SPR > 1) {
#pragma unroll
        for (int v = 0; v < VEC; ++v) acc[v] = subwarp_sum(acc[v], LPN, LPR);
      }
      if (writer) {
        Vec<T, VEC> xi;
        if constexpr (VEC == 1) xi.v[0] = xs[r * CW + cl];
        else xi = *reinterpret_cast<const Vec<T, VEC>*>(xs + r * CW + cl * VEC);   // already scaled by pre
        const T d = dgs[r];
        const T po = pss[r];
        Vec<T, VEC> out;
#pragma unroll
        for (int v = 0; v < VEC; ++v) out.v[v] = po * (d * xi.v[v] - acc[v]);
        st_vec<T, VEC>(g.y + yrow * g.ldy + cbase, out);
        if (g.dot_out) {
          if (g.dot_is_x) dw = xi;
#pragma unroll
          for (int v = 0; v < VEC; ++v) dsum[v] = fma(dw.v[v], out.v[v], dsum[v]);
        }
      }
    }
    __syncthreads();   // the next tile's phase A overwrites the staging buffers
  }

  if (g.dot_out) spmm_dot_epilogue<T, VEC, LPN, CW, kTiledBlock>(dsum, g.cw, g.c0, g.partials, g.counter, g.dot_out);
}

template <typename T, int VEC, int LPN, int LPR, int R>
static size_t tiled_smem_bytes(int lmax, int nzcap) {
  return (size_t)lmax * LPN * VEC * sizeof(T) + (size_t)nzcap * (sizeof(T) + 2) + (size_t)(R + 4) * 4 + (size_t)2 * R * sizeof(T) + 16;
}

constexpr size_t kTiledSmemLimit = 200 * 1024;

template <typename T, int VEC, int LPN, int LPR, int R>
static int launch_tiled(const TiledArgs<T>& g, cudaStream_t st) {
  const size_t smem = tiled_smem_bytes<T, VEC, LPN, LPR, R>(g.lmax, g.nzcap);
  if (smem > kTiledSmemLimit) {
    set_error("lap_spmm_tiled: tile needs %zu bytes of shared memory (> %zu): graph has no locality, use the CSR kernel",
              smem, kTiledSmemLimit);
    return MGP_EUNSUPPORTED;
  }
  auto kern = lap_spmm_tiled_kernel<T, VEC, LPN, LPR, R>;
  static size_t configured = 0;   // per instantiation
  if (smem > configured) {
    MGP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTiledSmemLimit));
    configured = kTiledSmemLimit;
  }
  int per_sm = (int)((220 * 1024) / (smem + 1024));
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 8) per_sm = 8;
  int64_t blocks = (int64_t)kNumSMs * per_sm;
  if (blocks > g.ntiles) blocks = g.ntiles;
  kern<<<(unsigned)blocks, kTiledBlock, smem, st>>>(g);
  MGP_LAUNCH_CHECK();
  return MGP_OK;
}

template <typename T>
static int lap_spmm_tiled(const int* rowptr, const unsigned short* lcol, const T* a, const T* diag, const int* halo_ptr,
                          const int* halo_col, int tile_rows, int lmax, int nzmax, const T* shift, const T* pre,
                          const T* post, const int* xmap, const int* ymap, const T* x, int64_t ldx, T* y, int64_t ldy,
                          int64_t n, int ncols, const T* dot_with, T* dot_out, void* dot_ws, cudaStream_t st) {
  constexpr int R = 128;
  MGP_CHECK_ARG(rowptr && lcol && a && diag && halo_ptr && halo_col && x && y, "lap_spmm_tiled: null pointer");
  MGP_CHECK_ARG(tile_rows == R, "lap_spmm_tiled: this build supports tile_rows == %d (got %d)", R, tile_rows);
  MGP_CHECK_ARG(n > 0 && ncols > 0 && ldx >= ncols && ldy >= ncols, "lap_spmm_tiled: bad shape");
  MGP_CHECK_ARG(x != y, "lap_spmm_tiled: X and Y must not alias");
  MGP_CHECK_ARG((dot_out == nullptr) || (dot_with && dot_ws), "lap_spmm_tiled: dot epilogue needs dot_with and dot_ws");
  MGP_CHECK_ARG(lmax >= R && lmax <= 65535 && nzmax > 0, "lap_spmm_tiled: bad tile statistics lmax=%d nzmax=%d", lmax, nzmax);
  constexpr int VECW = sizeof(T) == 4 ? 4 : 2;
  const bool aligned = (ldx % VECW == 0) && (ldy % VECW == 0) && (((uintptr_t)x) % 16 == 0) && (((uintptr_t)y) % 16 == 0) &&
                       (dot_with == nullptr || ((uintptr_t)dot_with) % 16 == 0);
  TiledArgs<T> g;
  g.rowptr = rowptr; g.lcol = lcol; g.a = a; g.diag = diag; g.halo_ptr = halo_ptr; g.halo_col = halo_col;
  g.shift = shift; g.pre = pre; g.post = post; g.xmap = xmap; g.ymap = ymap; g.x = x; g.ldx = ldx; g.y = y; g.ldy = ldy;
  g.n = n; g.ntiles = (int)ceil_div(n, R); g.lmax = (lmax + 3) & ~3; g.nzcap = (nzmax + 8 + 7) & ~7;
  g.dot_with = dot_out ? dot_with : nullptr; g.dot_out = dot_out;
  g.counter = dot_out ? reinterpret_cast<unsigned int*>(dot_ws) : nullptr;
  g.partials = dot_out ? reinterpret_cast<T*>(reinterpret_cast<char*>(dot_ws) + 256) : nullptr;
  g.dot_is_x = (dot_out && dot_with == x && pre == nullptr) ? 1 : 0;
  { const char* dbg = getenv("MGP_TILED_DEBUG"); g.debug = dbg ? atoi(dbg) : 0; }
  int c0 = 0;
  while (c0 < ncols) {
    const int rem = ncols - c0;
    int rc;
    g.c0 = c0;
    if (aligned && c0 % VECW == 0 && rem >= VECW) {
      if constexpr (sizeof(T) == 4) {
        if (rem >= 16) { g.cw = 16; rc = launch_tiled<T, 4, 4, 4, R>(g, st); }
        else if (rem >= 8) { g.cw = 8; rc = launch_tiled<T, 4, 2, 4, R>(g, st); }
        else { g.cw = 4; rc = launch_tiled<T, 4, 1, 4, R>(g, st); }
      } else {
        if (rem >= 16) { g.cw = 16; rc = launch_tiled<T, 2, 8, 32, R>(g, st); }
        else if (rem >= 8) { g.cw = 8; rc = launch_tiled<T, 2, 4, 16, R>(g, st); }
        else if (rem >= 4) { g.cw = 4; rc = launch_tiled<T, 2, 2, 8, R>(g, st); }
        else { g.cw = 2; rc = launch_tiled<T, 2, 1, 8, R>(g, st); }
      }
    } else {
      const int cw = rem > 32 ? 32 : rem;
      g.cw = cw;
      if (cw == 1) rc = launch_tiled<T, 1, 1, 4, R>(g, st);
      else if (cw == 2) rc = launch_tiled<T, 1, 2, 4, R>(g, st);
      else if (cw <= 4) rc = launch_tiled<T, 1, 4, 16, R>(g, st);
      else if (cw <= 8) rc = launch_tiled<T, 1, 8, 32, R>(g, st);
      else if (cw <= 16) rc = launch_tiled<T, 1, 16, 32, R>(g, st);
      else rc = launch_tiled<T, 1, 32, 32, R>(g, st);
    }
    if (rc != MGP_OK) return rc;
    c0 += g.cw;
  }
  return MGP_OK;
}

}  // namespace mgp

extern "C" {

int mgp_lap_spmm_tiled_f32(const int32_t* rowptr, const uint16_t* lcol, const float* a, const float* diag,
                           const int32_t* halo_ptr, const int32_t* halo_col, int32_t tile_rows, int32_t lmax, int32_t nzmax,
                           const float* shift, const float* pre, const float* post, const int32_t* xmap,
                           const int32_t* ymap, const float* x, int64_t ldx, float* y, int64_t ldy, int64_t n,
                           int32_t ncols, const float* dot_with, float* dot_out, void* dot_ws, void* stream) {
  return mgp::lap_spmm_tiled<float>(rowptr, lcol, a, diag, halo_ptr, halo_col, tile_rows, lmax, nzmax, shift, pre, post,
                                    xmap, ymap, x, ldx, y, ldy, n, ncols, dot_with, dot_out, dot_ws, (cudaStream_t)stream);
}

int mgp_lap_spmm_tiled_f64(const int32_t* rowptr, const uint16_t* lcol, const double* a, const double* diag,
                           const int32_t* halo_ptr, const int32_t* halo_col, int32_t tile_rows, int32_t lmax, int32_t nzmax,
                           const double* shift, const double* pre, const double* post, const int32_t* xmap,
                           const int32_t* ymap, const double* x, int64_t ldx, double* y, int64_t ldy, int64_t n,
                           int32_t ncols, const double* dot_with, double* dot_out, void* dot_ws, void* stream) {
  return mgp::lap_spmm_tiled<double>(rowptr, lcol, a, diag, halo_ptr, halo_col, tile_rows, lmax, nzmax, shift, pre, post,
                                     xmap, ymap, x, ldx, y, ldy, n, ncols, dot_with, dot_out, dot_ws, (cudaStream_t)stream);
}

}  // extern "C"
