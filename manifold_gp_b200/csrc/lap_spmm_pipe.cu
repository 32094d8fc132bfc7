// v3/v4 SpMM: the tile-compacted kernel of lap_spmm_tiled.cu as a warp-specialised, multi-stage TMA pipeline.
//
// v4 (this file): the (index, value) streams are stored PADDED -- every row holds a multiple of 4 entries (padding entries
// point at the row itself with weight 0) and every tile starts at a multiple of 8 entries -- so a row slot fetches 4
// indices with one 64-bit and 4 values with one 128-bit shared-memory load instead of 8 scalar loads: per 4 nonzeros
// the load/store pipe sees 2 + 16 wavefronts instead of 8 + 16 and the issue slots ~34 instead of ~56 instructions.
//
//   Y = post .* ( (diag + shift) .* X  -  A X )            (no `pre` scaling on this path; the host routes `pre` to v2)
//
// v2 measurements (ncu, profiles/): staging a tile (phase A, latency bound) and walking its rows (phase B, shared-
// memory wavefront bound) cost about the same and simply add up, because both phases queue on the same load/store pipe
// and every tile begins with a chain of dependent global loads.  Here a dedicated producer warp moves each tile with
// the bulk-copy engine (cp.async.bulk -> UBLKCP: index stream, value stream, the tile's own X rows and one 64-byte
// copy per halo row, all completing on an mbarrier transaction count) into one of two shared-memory stages while the
// consumer warps walk the rows of the other stage.  The load/store pipe then only sees the phase-B shared loads.
//
//   producer (warp 0):  wait empty[s] -> plain loads of row offsets / diagonal -> expect_tx + bulk copies -> arrive full[s]
//   consumers (16 warps): wait full[s] -> rows (one 4-lane slot per row for 16 fp32 columns) -> Y stores -> arrive empty[s]
#include "common.cuh"
#include "spmm_common.cuh"
#include "pipe_common.cuh"

namespace mgp {

constexpr int kPipeMaxStages = 3;
constexpr int kPipeProducerWarps = 4;
constexpr int kPipeProducerThreads = kPipeProducerWarps * 32;
constexpr int kPipeConsumerWarps = 16;
constexpr int kPipeThreads = (kPipeConsumerWarps + kPipeProducerWarps) * 32;

template <typename T>
struct PipeArgs {
  const int* rowptr;
  const unsigned short* lcol;
  const T* a;
  const T* diag;
  const int* halo_ptr;
  const int* halo_col;
  const T* shift;
  const T* post;
  const int* xmap;
  const int* ymap;
  const T* x;
  int64_t ldx;
  T* y;
  int64_t ldy;
  int64_t n;
  int ntiles;
  int lmax;
  int nzcap;
  int c0, cw;
  const T* dot_with;
  T* dot_out;
  T* partials;
  unsigned int* counter;
  int dot_is_x;
  int stages;   // 2 or 3 shared-memory stages (as many as fit)
};

// 4 consecutive values of the (padded, 16-byte aligned) value stream: unit u = entries [4u, 4u + 4)
template <typename T>
__device__ __forceinline__ void load_vals4(const T* vs, int u, T (&w)[4]) {
  if constexpr (sizeof(T) == 4) {
    const float4 t = reinterpret_cast<const float4*>(vs)[u];
    w[0] = t.x; w[1] = t.y; w[2] = t.z; w[3] = t.w;
  } else {
    const double2 t0 = reinterpret_cast<const double2*>(vs)[2 * u];
    const double2 t1 = reinterpret_cast<const double2*>(vs)[2 * u + 1];
    w[0] = t0.x; w[1] = t0.y; w[2] = t1.x; w[3] = t1.y;
  }
}

template <typename T, int CW, int R>
__host__ __device__ inline size_t pipe_stage_bytes(int lmax, int nzcap) {
  // xs [lmax][CW] T | vs [nzcap] T | cs [nzcap] u16 | dg [R] T | rp [R + 4] int   (every piece a multiple of 16 bytes)
  return (size_t)lmax * CW * sizeof(T) + (size_t)nzcap * (sizeof(T) + 2) + (size_t)R * sizeof(T) + (size_t)(R + 4) * 4;
}

template <typename T, int VEC, int LPN, int R>
__global__ void __launch_bounds__(kPipeThreads, 1)
lap_spmm_pipe_kernel(const PipeArgs<T> g) {
  constexpr int CW = LPN * VEC;
  constexpr int LPR = LPN;                       // one slot per row: no cross-lane reduction
  constexpr int ROWS_PER_WARP = 32 / LPR;
  constexpr int ROWS_PER_ROUND = kPipeConsumerWarps * ROWS_PER_WARP;
  constexpr uint32_t ROW_BYTES = CW * sizeof(T);
  static_assert(ROW_BYTES % 16 == 0, "bulk copies move multiples of 16 bytes");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kPipeMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kPipeMaxStages];
  const int kPipeStages = g.stages;

  const size_t stage_bytes = pipe_stage_bytes<T, CW, R>(g.lmax, g.nzcap);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  if (tid == 0) {
    for (int s = 0; s < kPipeStages; ++s) {
      mbar_init(&full_bar[s], 2 * kPipeProducerThreads);   // per producer thread: one plain arrive (thread 0: expect_tx) + one cp.async arrive
      mbar_init(&empty_bar[s], kPipeConsumerWarps);   // one elected lane per consumer warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  auto stage_ptrs = [&](int s, T*& xs, T*& vs, unsigned short*& cs, T*& dg, int*& rp) {
    unsigned char* b = smem_raw + (size_t)s * stage_bytes;
    xs = reinterpret_cast<T*>(b);
    vs = xs + (size_t)g.lmax * CW;
    cs = reinterpret_cast<unsigned short*>(vs + g.nzcap);
    dg = reinterpret_cast<T*>(cs + g.nzcap);
    rp = reinterpret_cast<int*>(dg + R);
  };

  T dsum[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) dsum[v] = T(0);

  if (warp < kPipeProducerWarps) {
    // =================================================== producer ===================================================
    // Tile metadata (row offsets, diagonal, halo ids) is pulled into L1 one tile ahead, so that after the stage frees up
    // the producer only issues copies: three large bulk copies (index stream, value stream, the tile's own X rows) on
    // the TMA engine and 16-byte cp.async for the scattered halo rows (per-row TMA descriptors would cost ~50-90 cycles
    // of TMA issue each and serialise the pipeline).
    constexpr int CH = ROW_BYTES / 16;               // 16-byte chunks per X row
    constexpr int MAXB = 5;                          // row batches per thread: lmax <= 640 rows / 128 threads
    struct Meta { int p0, p1, h0, nh; };
    auto load_meta = [&](int tile) {
      Meta m{0, 0, 0, 0};
      if (tile < g.ntiles) {
        const int64_t row0 = (int64_t)tile * R;
        const int nrows = (int)min((int64_t)R, g.n - row0);
        m.p0 = __ldg(g.rowptr + row0);
        m.p1 = __ldg(g.rowptr + row0 + nrows);
        m.h0 = __ldg(g.halo_ptr + tile);
        m.nh = __ldg(g.halo_ptr + tile + 1) - m.h0;
      }
      return m;
    };
    // Software pipeline inside the producer: while the copies of tile t are issued, the row ids / offsets / diagonal of
    // tile t+1 and the scalars of tile t+2 are already in flight, so no step waits on a global-load latency.
    struct Loads { int64_t src[MAXB]; int rpv[2]; T dgv; };
    const bool own_contig = (g.xmap == nullptr) && (g.ldx == CW);
    auto issue_loads = [&](int tile, const Meta& m) {
      Loads L;
#pragma unroll
      for (int b = 0; b < MAXB; ++b) L.src[b] = -1;
      L.rpv[0] = L.rpv[1] = 0;
      L.dgv = T(0);
      if (tile >= g.ntiles) return L;
      const int64_t row0 = (int64_t)tile * R;
      const int nrows = (int)min((int64_t)R, g.n - row0);
      const int base = m.p0 & ~7;
      const int nown = own_contig ? 0 : nrows;
      const int nscat = nown + m.nh;
#pragma unroll
      for (int b = 0; b < MAXB; ++b) {
        const int rr = tid + b * kPipeProducerThreads;
        if (rr < nscat) L.src[b] = rr < nown ? (int64_t)(row0 + rr) : (int64_t)__ldg(g.halo_col + m.h0 + rr - nown);
      }
      if (tid <= nrows) L.rpv[0] = __ldg(g.rowptr + row0 + tid) - base;
      if (tid + kPipeProducerThreads <= nrows) L.rpv[1] = __ldg(g.rowptr + row0 + tid + kPipeProducerThreads) - base;
      if (tid < nrows) L.dgv = __ldg(g.diag + row0 + tid);
      if (g.xmap) {
#pragma unroll
        for (int b = 0; b < MAXB; ++b) if (L.src[b] >= 0) L.src[b] = (int64_t)__ldg(g.xmap + L.src[b]);
      }
      return L;
    };
    Meta cur = load_meta(blockIdx.x);
    Meta nxt = load_meta(blockIdx.x + gridDim.x);
    Loads lcur = issue_loads(blockIdx.x, cur);
    int it = 0;
    for (int tile = blockIdx.x; tile < g.ntiles; tile += gridDim.x, ++it) {
      const int s = it % kPipeStages;
      const uint32_t ph = (it / kPipeStages) & 1;
      T *xs, *vs, *dg; unsigned short* cs; int* rp;
      stage_ptrs(s, xs, vs, cs, dg, rp);
      const int64_t row0 = (int64_t)tile * R;
      const int nrows = (int)min((int64_t)R, g.n - row0);
      const int base = cur.p0 & ~7;
      const int cnt8 = (cur.p1 - base + 7) & ~7;
      const int h0 = cur.h0, nh = cur.nh;
      const int nown = own_contig ? 0 : nrows;
      const int nscat = nown + nh;
      // loads for the following tiles go out first
      const Loads lnext = issue_loads(tile + gridDim.x, nxt);
      const Meta nxt2 = load_meta(tile + 2 * gridDim.x);
      mbar_wait(&empty_bar[s], ph ^ 1);                       // fresh barrier: parity 1 passes immediately
      if (tid == 0) {
        const uint32_t bytes = (uint32_t)cnt8 * (2 + (uint32_t)sizeof(T)) + (own_contig ? (uint32_t)nrows * ROW_BYTES : 0u);
        mbar_arrive_expect_tx(&full_bar[s], bytes);
        bulk_g2s(cs, g.lcol + base, (uint32_t)cnt8 * 2, &full_bar[s]);
        bulk_g2s(vs, g.a + base, (uint32_t)cnt8 * (uint32_t)sizeof(T), &full_bar[s]);
        if (own_contig) bulk_g2s(xs, g.x + row0 * g.ldx + g.c0, (uint32_t)nrows * ROW_BYTES, &full_bar[s]);
      }
      // scattered X rows: CH 16-byte cp.async per row
#pragma unroll
      for (int b = 0; b < MAXB; ++b) {
        const int rr = tid + b * kPipeProducerThreads;
        if (lcur.src[b] >= 0) {
          const int dstrow = rr < nown ? rr : R + (rr - nown);
          unsigned char* d = reinterpret_cast<unsigned char*>(xs + (size_t)dstrow * CW);
          const unsigned char* sp = reinterpret_cast<const unsigned char*>(g.x + lcur.src[b] * g.ldx + g.c0);
#pragma unroll
          for (int c = 0; c < CH; ++c) cp_async16(d + c * 16, sp + c * 16);
        }
      }
      for (int rr = tid + MAXB * kPipeProducerThreads; rr < nscat; rr += kPipeProducerThreads) {   // very large halos
        int64_t sr = rr < nown ? (int64_t)(row0 + rr) : (int64_t)__ldg(g.halo_col + h0 + rr - nown);
        if (g.xmap) sr = (int64_t)__ldg(g.xmap + sr);
        const int dstrow = rr < nown ? rr : R + (rr - nown);
        unsigned char* d = reinterpret_cast<unsigned char*>(xs + (size_t)dstrow * CW);
        const unsigned char* sp = reinterpret_cast<const unsigned char*>(g.x + sr * g.ldx + g.c0);
#pragma unroll
        for (int c = 0; c < CH; ++c) cp_async16(d + c * 16, sp + c * 16);
      }
      cp_async_arrive_noinc(&full_bar[s]);
      if (tid <= nrows) rp[tid] = lcur.rpv[0];
      if (tid + kPipeProducerThreads <= nrows) rp[tid + kPipeProducerThreads] = lcur.rpv[1];
      if (tid < nrows) dg[tid] = lcur.dgv;
      if (tid != 0) mbar_arrive(&full_bar[s]);
      cur = nxt; nxt = nxt2; lcur = lnext;
    }
  } else {
    // =================================================== consumers ==================================================
    const int cw_id = warp - kPipeProducerWarps;
    const int rg = lane / LPR;                        // row slot inside the warp
    const int cl = lane % LPN;
    const int cbase = g.c0 + cl * VEC;
    const T shift = g.shift ? *g.shift : T(0);
    int it = 0;
    for (int tile = blockIdx.x; tile < g.ntiles; tile += gridDim.x, ++it) {
      const int s = it % kPipeStages;
      const uint32_t ph = (it / kPipeStages) & 1;
      T *xs, *vs, *dg; unsigned short* cs; int* rp;
      stage_ptrs(s, xs, vs, cs, dg, rp);
      const int64_t row0 = (int64_t)tile * R;
      const int nrows = (int)min((int64_t)R, g.n - row0);
      mbar_wait(&full_bar[s], ph);
      for (int rb = 0; rb < R; rb += ROWS_PER_ROUND) {
        const int r = rb + cw_id * ROWS_PER_WARP + rg;
        const bool active = r < nrows;
        const int64_t row = row0 + r;
        int q0 = 0, q1 = 0;
        if (active) { q0 = rp[r]; q1 = rp[r + 1]; }
        Vec<T, VEC> dw;
        if (g.dot_out && !g.dot_is_x && active) {
          const int64_t drow = g.xmap ? (int64_t)__ldg(g.xmap + row) : row;
          dw = ldg_vec<T, VEC>(g.dot_with + drow * g.ldx + cbase);
        }
        int64_t yrow = row;
        if (g.ymap && active) yrow = (int64_t)__ldg(g.ymap + row);
        const T po = (g.post && active) ? __ldg(g.post + row) : T(1);
        T acc[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) acc[v] = T(0);
        // Entries are consumed 4 at a time (rows are padded to multiples of 4 entries): one 64-bit load brings 4 tile-local
        // indices, one 128-bit load (two for fp64) the 4 values.  The row slots of a warp advance in lockstep, so slot i
        // reads unit u_i + t at step t; rows of equal length start a fixed number of units apart and can sit in the same
        // banks.  Each slot therefore starts at a rotated unit (wrapping around at the end of the row): slot pairs
        // (2q, 2q+1) -- which share a quarter-warp and rely on the even/odd entry ordering for conflict-free X-row
        // loads -- keep a common rotation (one extra unit only if their rows are a multiple of 8 units apart), and the
        // leader of pair q starts 2q units (mod 8) beyond slot 0.
        const int ulen = (q1 - q0) >> 2;
        const int ubeg = q0 >> 2, uend = q1 >> 2;
        int rotu = 0;
        if constexpr (ROWS_PER_WARP >= 2) {
          const int u0 = __shfl_sync(0xffffffffu, ubeg, 0);
          const int lead_lane = (rg & ~1) * LPR;
          const int ul = __shfl_sync(0xffffffffu, ubeg, lead_lane);
          const int ulen_l = __shfl_sync(0xffffffffu, ulen, lead_lane);
          rotu = (2 * (rg >> 1) + u0 - ul) & 7;
          if ((rg & 1) && (ulen_l & 7) == 0) rotu += 1;
          if (rotu >= ulen) rotu = 0;
        }
        const uint2* cs2 = reinterpret_cast<const uint2*>(cs);
        const unsigned char* xb = reinterpret_cast<const unsigned char*>(xs) + cl * VEC * sizeof(T);
        int u = ubeg + rotu;
        uint2 jj = cs2[u];
        T w4[4];
        load_vals4<T>(vs, u, w4);
#pragma unroll 2
        for (int t = 0; t < ulen; ++t) {
          int un = u + 1;
          if (un == uend) un = ubeg;
          const uint2 jn = cs2[un];                   // next unit in flight while this one is consumed
          T wn[4];
          load_vals4<T>(vs, un, wn);
          const uint32_t o0 = (jj.x & 0xffffu) * ROW_BYTES, o1 = (jj.x >> 16) * ROW_BYTES;
          const uint32_t o2 = (jj.y & 0xffffu) * ROW_BYTES, o3 = (jj.y >> 16) * ROW_BYTES;
          const Vec<T, VEC> x0 = *reinterpret_cast<const Vec<T, VEC>*>(xb + o0);
          const Vec<T, VEC> x1 = *reinterpret_cast<const Vec<T, VEC>*>(xb + o1);
          const Vec<T, VEC> x2 = *reinterpret_cast<const Vec<T, VEC>*>(xb + o2);
          const Vec<T, VEC> x3 = *reinterpret_cast<const Vec<T, VEC>*>(xb + o3);
#pragma unroll
          for (int v = 0; v < VEC; ++v) acc[v] = fma(w4[0], x0.v[v], acc[v]);
#pragma unroll
          for (int v = 0; v < VEC; ++v) acc[v] = fma(w4[1], x1.v[v], acc[v]);
#pragma unroll
          for (int v = 0; v < VEC; ++v) acc[v] = fma(w4[2], x2.v[v], acc[v]);
#pragma unroll
          for (int v = 0; v < VEC; ++v) acc[v] = fma(w4[3], x3.v[v], acc[v]);
          jj = jn;
#pragma unroll
          for (int e = 0; e < 4; ++e) w4[e] = wn[e];
          u = un;
        }
        if (active) {
          const Vec<T, VEC> xi = *reinterpret_cast<const Vec<T, VEC>*>(xs + r * CW + cl * VEC);
          const T d = dg[r] + shift;
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
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[s]);      // this warp is done reading stage s
    }
  }

  if (g.dot_out) {
    __syncthreads();
    spmm_dot_epilogue<T, VEC, LPN, CW, kPipeThreads>(dsum, g.cw, g.c0, g.partials, g.counter, g.dot_out);
  }
}

constexpr size_t kPipeSmemLimit = 208 * 1024;   // dynamic; the kernel also has < 12 KB of static shared memory

template <typename T, int VEC, int LPN, int R>
static int launch_pipe(const PipeArgs<T>& g, cudaStream_t st) {
  constexpr int CW = LPN * VEC;
  const size_t one = pipe_stage_bytes<T, CW, R>(g.lmax, g.nzcap);
  PipeArgs<T> gg = g;
  gg.stages = (3 * one <= kPipeSmemLimit) ? 3 : 2;
  const size_t smem = gg.stages * one;
  if (smem > kPipeSmemLimit) return MGP_EUNSUPPORTED;
  auto kern = lap_spmm_pipe_kernel<T, VEC, LPN, R>;
  static size_t configured = 0;
  if (smem > configured) {
    MGP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  int64_t blocks = kNumSMs;
  if (blocks > g.ntiles) blocks = g.ntiles;
  kern<<<(unsigned)blocks, kPipeThreads, smem, st>>>(gg);
  MGP_LAUNCH_CHECK();
  return MGP_OK;
}

// One column pass of the pipelined kernel; returns MGP_EUNSUPPORTED when the pass does not qualify (caller falls back
// to the v2 kernel): needs VEC-wide aligned columns and two stages that fit in shared memory.
template <typename T>
static int lap_spmm_pipe_pass(const int* rowptr, const unsigned short* lcol, const T* a, const T* diag, const int* halo_ptr,
                       const int* halo_col, int lmax, int nzcap, const T* shift, const T* post, const int* xmap,
                       const int* ymap, const T* x, int64_t ldx, T* y, int64_t ldy, int64_t n, int c0, int cw,
                       const T* dot_with, T* dot_out, T* partials, unsigned int* counter, int dot_is_x, cudaStream_t st) {
  constexpr int R = 128;
  PipeArgs<T> g;
  g.rowptr = rowptr; g.lcol = lcol; g.a = a; g.diag = diag; g.halo_ptr = halo_ptr; g.halo_col = halo_col; g.shift = shift;
  g.post = post; g.xmap = xmap; g.ymap = ymap; g.x = x; g.ldx = ldx; g.y = y; g.ldy = ldy; g.n = n;
  g.ntiles = (int)ceil_div(n, R); g.lmax = lmax; g.nzcap = nzcap; g.c0 = c0; g.cw = cw; g.dot_with = dot_with;
  g.dot_out = dot_out; g.partials = partials; g.counter = counter; g.dot_is_x = dot_is_x;
  if constexpr (sizeof(T) == 4) {
    if (cw == 16) return launch_pipe<T, 4, 4, R>(g, st);
    if (cw == 8) return launch_pipe<T, 4, 2, R>(g, st);
    if (cw == 4) return launch_pipe<T, 4, 1, R>(g, st);
  } else {
    if (cw == 16) return launch_pipe<T, 2, 8, R>(g, st);
    if (cw == 8) return launch_pipe<T, 2, 4, R>(g, st);
    if (cw == 4) return launch_pipe<T, 2, 2, R>(g, st);
    if (cw == 2) return launch_pipe<T, 2, 1, R>(g, st);
  }
  return MGP_EUNSUPPORTED;
}

// ---- padded value stream: a_p[prowptr[r] + t] = a[rowptr[r] + t] for t < len(r), 0 for the padding entries -------------
template <typename T>
__global__ void __launch_bounds__(256)
lap_pad_values_kernel(const int* __restrict__ rowptr, const int* __restrict__ prowptr, const T* __restrict__ a, int64_t n,
                      T* __restrict__ ap) {
  constexpr int L = 8;   // lanes per row
  const int l = threadIdx.x & (L - 1);
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / L;
  if (row >= n) return;
  const int p0 = rowptr[row], len = rowptr[row + 1] - p0;
  const int q0 = prowptr[row], plen = prowptr[row + 1] - q0;
  for (int t = l; t < plen; t += L) ap[q0 + t] = t < len ? ld_stream(a + p0 + t) : T(0);
}

template <typename T>
static int lap_pad_values(const int* rowptr, const int* prowptr, const T* a, int64_t n, T* ap, cudaStream_t st) {
  MGP_CHECK_ARG(rowptr && prowptr && a && ap && n > 0, "lap_pad_values: bad arguments");
  const int64_t threads = n * 8;
  lap_pad_values_kernel<T><<<(unsigned)ceil_div(threads, (int64_t)256), 256, 0, st>>>(rowptr, prowptr, a, n, ap);
  MGP_LAUNCH_CHECK();
  return MGP_OK;
}

template <typename T>
static int lap_spmm_pipe(const int* prowptr, const unsigned short* plcol, const T* ap, const T* diag, const int* halo_ptr,
                         const int* halo_col, int tile_rows, int lmax, int pnzmax, const T* shift, const T* post,
                         const int* xmap, const int* ymap, const T* x, int64_t ldx, T* y, int64_t ldy, int64_t n, int ncols,
                         const T* dot_with, T* dot_out, void* dot_ws, cudaStream_t st) {
  constexpr int R = 128;
  constexpr int VECW = sizeof(T) == 4 ? 4 : 2;
  MGP_CHECK_ARG(prowptr && plcol && ap && diag && halo_ptr && halo_col && x && y, "lap_spmm_pipe: null pointer");
  MGP_CHECK_ARG(tile_rows == R, "lap_spmm_pipe: this build supports tile_rows == %d (got %d)", R, tile_rows);
  MGP_CHECK_ARG(n > 0 && ncols > 0 && ldx >= ncols && ldy >= ncols, "lap_spmm_pipe: bad shape");
  MGP_CHECK_ARG(x != y, "lap_spmm_pipe: X and Y must not alias");
  MGP_CHECK_ARG((dot_out == nullptr) || (dot_with && dot_ws), "lap_spmm_pipe: dot epilogue needs dot_with and dot_ws");
  MGP_CHECK_ARG(lmax >= R && lmax <= 65535 && pnzmax > 0 && pnzmax % 8 == 0, "lap_spmm_pipe: bad tile statistics lmax=%d pnzmax=%d",
                lmax, pnzmax);
  const bool aligned = (ldx % VECW == 0) && (ldy % VECW == 0) && (((uintptr_t)x) % 16 == 0) && (((uintptr_t)y) % 16 == 0) &&
                       (dot_with == nullptr || ((uintptr_t)dot_with) % 16 == 0) && (ncols % VECW == 0);
  if (!aligned) return MGP_EUNSUPPORTED;
  const int lmax4 = (lmax + 3) & ~3;
  {  // the widest pass must fit two stages, otherwise nothing is launched and the caller uses the v2 kernel
    int cw = VECW;
    while (cw * 2 <= ncols && cw * 2 <= 16) cw *= 2;
    const size_t one = (size_t)lmax4 * cw * sizeof(T) + (size_t)pnzmax * (sizeof(T) + 2) + (size_t)R * sizeof(T) + (size_t)(R + 4) * 4;
    if (2 * one > kPipeSmemLimit) return MGP_EUNSUPPORTED;
  }
  T* partials = dot_out ? reinterpret_cast<T*>(reinterpret_cast<char*>(dot_ws) + 256) : nullptr;
  unsigned int* counter = dot_out ? reinterpret_cast<unsigned int*>(dot_ws) : nullptr;
  const int dot_is_x = (dot_out && dot_with == x) ? 1 : 0;
  int c0 = 0;
  while (c0 < ncols) {
    const int rem = ncols - c0;
    int cw = VECW;
    while (cw * 2 <= rem && cw * 2 <= 16) cw *= 2;
    const int rc = lap_spmm_pipe_pass<T>(prowptr, plcol, ap, diag, halo_ptr, halo_col, lmax4, pnzmax, shift, post, xmap, ymap, x,
                                         ldx, y, ldy, n, c0, cw, dot_out ? dot_with : nullptr, dot_out, partials, counter,
                                         dot_is_x, st);
    if (rc != MGP_OK) return rc;
    c0 += cw;
  }
  return MGP_OK;
}

}  // namespace mgp

extern "C" {

int mgp_lap_pad_values_f32(const int32_t* rowptr, const int32_t* prowptr, const float* a, int64_t n, float* ap, void* stream) {
  return mgp::lap_pad_values<float>(rowptr, prowptr, a, n, ap, (cudaStream_t)stream);
}
int mgp_lap_pad_values_f64(const int32_t* rowptr, const int32_t* prowptr, const double* a, int64_t n, double* ap, void* stream) {
  return mgp::lap_pad_values<double>(rowptr, prowptr, a, n, ap, (cudaStream_t)stream);
}

int mgp_lap_spmm_pipe_f32(const int32_t* prowptr, const uint16_t* plcol, const float* ap, const float* diag,
                          const int32_t* halo_ptr, const int32_t* halo_col, int32_t tile_rows, int32_t lmax, int32_t pnzmax,
                          const float* shift, const float* post, const int32_t* xmap, const int32_t* ymap, const float* x,
                          int64_t ldx, float* y, int64_t ldy, int64_t n, int32_t ncols, const float* dot_with, float* dot_out,
                          void* dot_ws, void* stream) {
  return mgp::lap_spmm_pipe<float>(prowptr, plcol, ap, diag, halo_ptr, halo_col, tile_rows, lmax, pnzmax, shift, post, xmap,
                                   ymap, x, ldx, y, ldy, n, ncols, dot_with, dot_out, dot_ws, (cudaStream_t)stream);
}
int mgp_lap_spmm_pipe_f64(const int32_t* prowptr, const uint16_t* plcol, const double* ap, const double* diag,
                          const int32_t* halo_ptr, const int32_t* halo_col, int32_t tile_rows, int32_t lmax, int32_t pnzmax,
                          const double* shift, const double* post, const int32_t* xmap, const int32_t* ymap, const double* x,
                          int64_t ldx, double* y, int64_t ldy, int64_t n, int32_t ncols, const double* dot_with,
                          double* dot_out, void* dot_ws, void* stream) {
  return mgp::lap_spmm_pipe<double>(prowptr, plcol, ap, diag, halo_ptr, halo_col, tile_rows, lmax, pnzmax, shift, post, xmap,
                                    ymap, x, ldx, y, ldy, n, ncols, dot_with, dot_out, dot_ws, (cudaStream_t)stream);
}

}  // extern "C"
