// EXPERIMENTAL (not dispatched by default, NOT yet run on a GPU): the quad-row walk of lap_spmm_quad.cu inside the pipelined
// producer of lap_spmm_wi.cu.  Generated from lap_spmm_wi.cu at the end of round 1: the producer warps, stages, metadata / halo-id
// rings and barrier protocol are unchanged; a stage now holds [X rows | 4-wide value slots | 16-bit tile-local indices | warp
// offsets] of the quad streams (graph.quad_streams) and the 16 consumer warps each walk the union columns of two row quads (one
// X-row load per union column serves four matrix rows).  Same operation and reference lines as lap_spmm_wi.cu
// (graph_laplacian_operator.py:108-124 / precision_matern_operator.py:28-32).  Peer-memory mode is not wired.
// Opt-in test: MGP_TEST_EXPERIMENTAL=1 pytest tests/test_gpu_spmm_kernels.py -k quadpipe ; timing: BENCH_KERNELS=quadpipe.
#include "common.cuh"
#include "pipe_common.cuh"
#include "spmm_common.cuh"

namespace mgp {

constexpr int kQpMaxStages = 3;
constexpr int kQpConsumerWarps = 16;
constexpr int kQpRows = 128;                       // rows per tile = 16 consumer warps x 8 row slots
constexpr int kQpIdSlots = 4;                      // halo-id ring: ids are requested this many tiles ahead
constexpr int kQpChunk = 32;                       // tiles per metadata chunk
constexpr int kQpMetaW = 16 * kQpChunk + 4;        // ints of wptr per chunk (513 used)
constexpr int kQpMetaH = kQpChunk + 4;             // ints of hptr per chunk (33 used)
constexpr int kQpMetaBytes = (kQpMetaW + kQpMetaH) * 4;
constexpr size_t kQpSmemLimit = 232448 - 10240;   // dynamic; static shared memory (dot epilogue, barriers) stays < 8 KB

template <typename T>
struct QpArgs {
  const int* wptr;               // qwptr [16 * ntiles + 1] stream offsets of the per-warp blocks (multiples of 8 entries); padded like wptr
  const int* qrows;              // [32 * ntiles][4] rows of every quad (-1 = none)
  const unsigned short* wcol;    // qidx: tile-local column per stream entry
  const T* aw;                   // qval: 4 values per stream entry (the quad's rows; 0 where absent / padding)
  const T* diag;
  const int* hptr;               // [ntiles + 1] halo list offsets (multiples of 4); padded, see header
  const int* hcol;               // halo row ids, every tile's list padded to a multiple of 4 with valid ids
  const T* shift;
  const T* post;
  const int* xmap;
  const int* ymap;
  const T* x;
  const unsigned char* const* peer_x;   // multi-GPU: device array of the ranks' X base pointers (peer-mapped); halo ids are then
                                        // (rank << 26) | row-in-that-rank's-X and the halo rows are fetched over NVLink.  NULL: ids index x.
  int npeers;
  int rank;
  unsigned int* const* sync_flags;     // optional fused cross-GPU barrier: device array of the ranks' flag arrays uint32[npeers]
  const T* sync_epoch;                 // epoch = (unsigned)*sync_epoch + 1 (the CG iteration counter): block 0 publishes it to
                                        // every peer at kernel start, producers wait for all peers' flags before the first remote row
  int64_t ldx;
  T* y;
  int64_t ldy;
  int64_t n;
  int ntiles;
  int lmax;      // max rows of X a tile stages (own + padded halo), multiple of 4
  int nzcap;     // max stream entries of a tile + 32 (multiple of 32)
  int hmax;      // max padded halo length (multiple of 4)
  int c0;
  const T* dot_with;
  T* dot_out;
  T* partials;
  unsigned int* counter;
  int dot_is_x;
  int stages;
  int debug;     // timing experiments only (MGP_WI_DEBUG bit mask): 1 = consumers skip the row walk, 2 = producers skip the halo rows
                 // (same-process A/B on B200, cfg-C: 150 us full, 104 without halo copies, 90 without the walk, 62 with neither)
};

template <typename T>
__host__ __device__ inline size_t qp_stage_bytes(int lmax, int nzcap) {
  // xs [lmax] x 64 B | vs [nzcap] x 4 T | cs [nzcap] u16 | rp [20] int       (every piece a multiple of 16 bytes)
  return (size_t)lmax * 64 + (size_t)nzcap * (4 * sizeof(T) + 2) + 20 * 4;
}
__host__ __device__ inline size_t qp_ring_bytes(int hmax) { return 2 * (size_t)kQpMetaBytes + (size_t)kQpIdSlots * hmax * 4; }

template <int PW>
__device__ __forceinline__ void qp_producers_sync() { asm volatile("bar.sync 1, %0;" ::"n"(PW * 32) : "memory"); }

// Design notes (what three rounds of ncu / timing experiments on B200 established, profiles/):
//  * consumers are bound by shared-memory wavefronts: 16 per 32 nonzeros for the 64-byte X rows; the warp-interleaved
//    streams bring the (index, value) part down from 6-8 to 2;
//  * the PRODUCER decides whether they ever get there.  With the tile's metadata, row offsets, diagonal and halo ids
//    loaded from global memory into registers (v3-v5a) each producer warp ran ~500 dependent instructions per tile and
//    waited on global-load latency every iteration (register rings do not help: a scoreboard wait covers every load in
//    flight on that scoreboard); the kernel took the same ~105 us with the consumers' row walk switched off.  A bare
//    TMA ring with the same barrier protocol streams at 7 TB/s (profiles/micro/tma_stream.cu), so the fix is to keep
//    long-latency loads out of the producer altogether:
//      - a block owns a CONTIGUOUS range of tiles, so its metadata is contiguous: chunks of 32 tiles of (wptr, hptr) are
//        bulk-copied into a 2-slot shared-memory ring one chunk ahead;
//      - each tile's halo id list is bulk-copied into a 4-slot ring 4 tiles ahead;
//      - the producers then only read shared memory (tens of cycles) and issue copies: thread 0 the three bulk copies of
//        the stage, thread 32 the ring refills, all 128 the 16-byte cp.async of the halo rows (4 lanes per row);
//      - the diagonal is read by the consumers themselves (issued before the row walk, used after it).
template <typename T, int PW>
__global__ void __launch_bounds__((kQpConsumerWarps + PW) * 32, 1)
lap_spmm_qp_kernel(const QpArgs<T> g) {
  constexpr int kQpProducerWarps = PW, kQpProducerThreads = PW * 32, kQpThreads = (kQpConsumerWarps + PW) * 32;
  constexpr int R = kQpRows;
  constexpr int VEC = 16 / sizeof(T);          // elements per 16-byte chunk
  constexpr int CW = 4 * VEC;                  // columns per pass: 64-byte rows
  constexpr uint32_t ROW_BYTES = 64;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kQpMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kQpMaxStages];
  __shared__ __align__(8) uint64_t meta_bar[2];
  __shared__ __align__(8) uint64_t ids_bar[kQpIdSlots];
  __shared__ const unsigned char* peer_tab[32];
  const int nstages = g.stages;
  const size_t stage_bytes = qp_stage_bytes<T>(g.lmax, g.nzcap);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  unsigned char* const ring = smem_raw + (size_t)nstages * stage_bytes;
  int* const ids_ring = reinterpret_cast<int*>(ring + 2 * kQpMetaBytes);

  if (g.peer_x && tid < g.npeers) peer_tab[tid] = g.peer_x[tid];
  if (tid == 0) {
    for (int s = 0; s < nstages; ++s) {
      mbar_init(&full_bar[s], 2 * kQpProducerThreads);   // per producer thread: one plain arrive (thread 0: expect_tx) + one cp.async arrive
      mbar_init(&empty_bar[s], 16);                      // one arrival per warp block of the tile
    }
    mbar_init(&meta_bar[0], 1);
    mbar_init(&meta_bar[1], 1);
    for (int s = 0; s < kQpIdSlots; ++s) mbar_init(&ids_bar[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  unsigned int sync_epoch = 0u;
  if (g.sync_flags) {
    sync_epoch = (unsigned int)(*g.sync_epoch) + 1u;
    if (blockIdx.x == 0 && warp == 1 && lane < g.npeers) {      // "everything enqueued before this launch is done on my side"
      __threadfence_system();
      st_release_sys(g.sync_flags[lane] + g.rank, sync_epoch);
    }
  }

  // contiguous tile range of this block
  const int t0 = (int)(((int64_t)blockIdx.x * g.ntiles) / gridDim.x);
  const int t1 = (int)(((int64_t)(blockIdx.x + 1) * g.ntiles) / gridDim.x);

  T dsum[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) dsum[v] = T(0);

  if (warp < kQpProducerWarps) {
    // =================================================== producer ===================================================
    const int pw = warp, sub = lane >> 2, ch = lane & 3;          // producer warp / row of a pass / 16-byte chunk
    const int ch0 = t0 >> 5, ch_last = (t1 - 1) >> 5;             // metadata chunks this block touches
    const bool own_contig = (g.xmap == nullptr) && (g.ldx == CW);
    const unsigned char* const xbase = reinterpret_cast<const unsigned char*>(g.x + g.c0) + ch * 16;
    const int64_t ldxb = g.ldx * (int64_t)sizeof(T);
    const int64_t xoff = (int64_t)g.c0 * (int64_t)sizeof(T) + ch * 16;   // same column window / chunk in a peer's X
    auto meta_w = [&](int c) { return reinterpret_cast<const int*>(ring + (size_t)((c - ch0) & 1) * kQpMetaBytes); };
    auto meta_h = [&](int c) { return meta_w(c) + kQpMetaW; };
    auto request_meta = [&](int c) {                               // one thread
      uint64_t* bar = &meta_bar[(c - ch0) & 1];
      mbar_arrive_expect_tx(bar, (uint32_t)kQpMetaBytes);
      bulk_g2s(const_cast<int*>(meta_w(c)), g.wptr + (size_t)c * 16 * kQpChunk, kQpMetaW * 4, bar);
      bulk_g2s(const_cast<int*>(meta_h(c)), g.hptr + (size_t)c * kQpChunk, kQpMetaH * 4, bar);
    };
    auto wait_meta = [&](int c) { mbar_wait(&meta_bar[(c - ch0) & 1], (uint32_t)(((c - ch0) >> 1) & 1)); };
    auto request_ids = [&](int t) {                                // one thread; the chunk of tile t must have landed
      const int i = t - t0;
      const int* hp = meta_h(t >> 5) + (t & 31);
      const int h0 = hp[0], nh = hp[1] - h0;
      uint64_t* bar = &ids_bar[i % kQpIdSlots];
      mbar_arrive_expect_tx(bar, (uint32_t)nh * 4u);
      if (nh > 0) bulk_g2s(ids_ring + (size_t)(i % kQpIdSlots) * g.hmax, g.hcol + h0, (uint32_t)nh * 4u, bar);
    };
    if (tid == 32 && t0 < t1) {                                    // prologue of the two rings
      request_meta(ch0);
      if (ch0 < ch_last) request_meta(ch0 + 1);
      wait_meta(ch0);
      for (int t = t0; t < t1 && t < t0 + kQpIdSlots; ++t) {
        if ((t >> 5) != ch0) wait_meta(t >> 5);
        request_ids(t);
      }
    }
    int s = 0;
    uint32_t ph = 0;
    bool peers_ready = false;
    for (int t = t0; t < t1; ++t) {
      const int i = t - t0;
      unsigned char* const sb = smem_raw + (size_t)s * stage_bytes;
      unsigned char* const xs = sb;
      T* const vs = reinterpret_cast<T*>(sb + (size_t)g.lmax * ROW_BYTES);
      unsigned short* const cs = reinterpret_cast<unsigned short*>(vs + 4 * (size_t)g.nzcap);
      int* const rp = reinterpret_cast<int*>(cs + g.nzcap);
      const int c = t >> 5;
      wait_meta(c);                                                // passes at once except on the first tile of a chunk
      const int* wp = meta_w(c) + 16 * (t & 31);
      const int* hp = meta_h(c) + (t & 31);
      const int base = wp[0];
      const int cnt = wp[16] - base;                               // multiple of 8 entries
      const int nh = hp[1] - hp[0];
      const int64_t row0 = (int64_t)t * R;
      const int nrows = (int)min((int64_t)R, g.n - row0);
      const int nown = own_contig ? 0 : nrows;
      const int nscat = (g.debug & 2) ? 0 : nown + nh;
      const int* ids = ids_ring + (size_t)(i % kQpIdSlots) * g.hmax;
      mbar_wait(&ids_bar[i % kQpIdSlots], (uint32_t)((i / kQpIdSlots) & 1));
      mbar_wait(&empty_bar[s], ph ^ 1);                            // fresh barrier: parity 1 passes immediately
      if (tid == 0) {
        const uint32_t bytes = (uint32_t)cnt * (2 + 4 * (uint32_t)sizeof(T)) + (own_contig ? (uint32_t)nrows * ROW_BYTES : 0u);
        mbar_arrive_expect_tx(&full_bar[s], bytes);
        if (cnt > 0) {
          bulk_g2s(cs, g.wcol + base, (uint32_t)cnt * 2, &full_bar[s]);
          bulk_g2s(vs, g.aw + (size_t)base * 4, (uint32_t)cnt * 4 * (uint32_t)sizeof(T), &full_bar[s]);
        }
        if (own_contig) bulk_g2s(xs, g.x + row0 * g.ldx + g.c0, (uint32_t)nrows * ROW_BYTES, &full_bar[s]);
      }
      // scattered X rows: 4 consecutive lanes copy the four 16-byte chunks of one row (8 whole rows per warp instruction).
      // Fused cross-GPU barrier: a warp waits for the peers' flags only when it reaches the first row that actually lives
      // on ANOTHER rank -- tiles away from the partition boundaries (most of them) never wait, so the NVLink round trip
      // of the barrier hides behind the interior tiles.
      for (int rb = 8 * pw; rb < nscat; rb += 8 * kQpProducerWarps) {
        const int rr = rb + sub;
        const bool valid = rr < nscat;
        int sr = 0;
        const unsigned char* xb = xbase;
        if (valid) {
          sr = rr < nown ? (int)(row0 + rr) : ids[rr - nown];
          if (g.xmap) sr = __ldg(g.xmap + sr);
        }
        bool remote = false;
        if (valid && g.peer_x && rr >= nown) {                     // halo row: read it from its owner's X (NVLink if remote)
          const int owner = sr >> 26;
          xb = peer_tab[owner] + xoff;
          sr &= (1 << 26) - 1;
          remote = owner != g.rank;
        }
        if (g.sync_flags && !peers_ready && __any_sync(0xffffffffu, remote)) {
          if (lane < g.npeers) {
            const unsigned int* f = g.sync_flags[g.rank] + lane;
            unsigned int spins = 0;
            while ((int)(ld_acquire_sys(f) - sync_epoch) < 0) {
              if (++spins > (1u << 25)) __trap();
            }
          }
          __syncwarp();
          peers_ready = true;
        }
        if (valid) {
          const int dstrow = rr < nown ? rr : R + (rr - nown);
          cp_async16(xs + (size_t)dstrow * ROW_BYTES + ch * 16, xb + (int64_t)sr * ldxb);
        }
      }
      cp_async_arrive_noinc(&full_bar[s]);
      if (tid <= 16) rp[tid] = wp[tid] - base;
      if (tid != 0) mbar_arrive(&full_bar[s]);
      qp_producers_sync<PW>();                                          // everyone is done with this tile's id slot and metadata
      if (tid == 32) {
        const int tn = t + kQpIdSlots;
        if (tn < t1) {
          if ((tn >> 5) != c) wait_meta(tn >> 5);
          request_ids(tn);
        }
        // last tile of a chunk done: its ring slot is free -> fetch the chunk after the next one into it
        if (((t + 1) & 31) == 0 && c + 2 <= ch_last) request_meta(c + 2);
      }
      if (++s == nstages) { s = 0; ph ^= 1; }
    }
  } else {
    // =================================================== consumers ==================================================
    const int grp = lane >> 2;                       // lane group: quad (grp >> 2) of the warp block, sub-list (grp & 3)
    const int ch = lane & 3;                         // 16-byte chunk of the 64-byte rows this lane owns
    const int sub = grp & 3;                         // ... and the row slot of the quad this lane group finishes
    const int cbase = g.c0 + ch * VEC;
    const T shift = g.shift ? *g.shift : T(0);
    for (unsigned int tk = (unsigned int)(warp - kQpProducerWarps);; tk += 16) {
      const int ti = (int)(tk >> 4);
      if (ti >= t1 - t0) break;
      const int w = (int)(tk & 15u);                 // warp block of the tile: quads 2w, 2w + 1
      const int tile = t0 + ti;
      int s;
      uint32_t ph;
      if (nstages == 3) { s = ti % 3; ph = (uint32_t)(ti / 3) & 1u; } else { s = ti & 1; ph = (uint32_t)(ti >> 1) & 1u; }
      unsigned char* const sb = smem_raw + (size_t)s * stage_bytes;
      unsigned char* const xs = sb;
      const T* const vs = reinterpret_cast<const T*>(sb + (size_t)g.lmax * ROW_BYTES);
      const unsigned short* const cs = reinterpret_cast<const unsigned short*>(vs + 4 * (size_t)g.nzcap);
      const int* const rp = reinterpret_cast<const int*>(cs + g.nzcap);
      const int64_t row0 = (int64_t)tile * R;
      // operands of the epilogue that live in global memory: in flight while the stage is awaited and walked
      const int64_t row = (int64_t)__ldg(g.qrows + (((size_t)tile * 16 + w) * 2 + (grp >> 2)) * 4 + sub);
      const bool active = row >= 0;
      Vec<T, VEC> dw;
      if (g.dot_out && !g.dot_is_x && active) {
        const int64_t drow = g.xmap ? (int64_t)__ldg(g.xmap + row) : row;
        dw = ldg_vec<T, VEC>(g.dot_with + drow * g.ldx + cbase);
      }
      int64_t yrow = row;
      if (g.ymap && active) yrow = (int64_t)__ldg(g.ymap + row);
      const T po = (g.post && active) ? __ldg(g.post + row) : T(1);
      const T dgv = active ? __ldg(g.diag + row) : T(0);
      mbar_wait(&full_bar[s], ph);
      const int ofs = rp[w];
      const int steps = (g.debug & 1) ? 0 : (rp[w + 1] - ofs) >> 3;
      const unsigned short* cp = cs + ofs + grp;
      const T* vp = vs + (size_t)(ofs + grp) * 4;
      T acc[4][VEC];
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int v = 0; v < VEC; ++v) acc[c][v] = T(0);
#pragma unroll 2
      for (int t = 0; t < steps; ++t) {
        const uint32_t j = cp[0];
        T q4[4];
        if constexpr (sizeof(T) == 4) {
          const float4 q = *reinterpret_cast<const float4*>(vp);
          q4[0] = q.x; q4[1] = q.y; q4[2] = q.z; q4[3] = q.w;
        } else {
          const double2 qa = *reinterpret_cast<const double2*>(vp), qb = *(reinterpret_cast<const double2*>(vp) + 1);
          q4[0] = qa.x; q4[1] = qa.y; q4[2] = qb.x; q4[3] = qb.y;
        }
        cp += 8;
        vp += 32;
        const Vec<T, VEC> xv = *reinterpret_cast<const Vec<T, VEC>*>(xs + j * ROW_BYTES + ch * 16);
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int v = 0; v < VEC; ++v) acc[c][v] = fma(q4[c], xv.v[v], acc[c][v]);
      }
      // sum the four sub-lists of a quad: lanes that differ in bits 2 and 3 (same quad, same chunk)
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
          acc[c][v] += __shfl_xor_sync(0xffffffffu, acc[c][v], 4);
          acc[c][v] += __shfl_xor_sync(0xffffffffu, acc[c][v], 8);
        }
      if (active) {
        T res[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) res[v] = sub == 0 ? acc[0][v] : sub == 1 ? acc[1][v] : sub == 2 ? acc[2][v] : acc[3][v];
        const int r = (int)(row - row0);
        const Vec<T, VEC> xi = *reinterpret_cast<const Vec<T, VEC>*>(xs + (size_t)r * ROW_BYTES + ch * 16);
        const T d = dgv + shift;
        Vec<T, VEC> out;
#pragma unroll
        for (int v = 0; v < VEC; ++v) out.v[v] = po * (d * xi.v[v] - res[v]);
        st_vec<T, VEC>(g.y + yrow * g.ldy + cbase, out);
        if (g.dot_out) {
          if (g.dot_is_x) dw = xi;
#pragma unroll
          for (int v = 0; v < VEC; ++v) dsum[v] = fma(dw.v[v], out.v[v], dsum[v]);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[s]);      // this block of stage s is done (16 arrivals free the stage)
    }
  }

  if (g.dot_out) {
    __syncthreads();
    spmm_dot_epilogue<T, VEC, 4, CW, kQpThreads>(dsum, CW, g.c0, g.partials, g.counter, g.dot_out);
  }
}


template <typename T>
static int lap_spmm_qp(const int* qwptr, const unsigned short* qidx, const T* qval, const int* qrows, const T* diag, const int* hptr,
                       const int* hcol, int tile_rows, int lmax, int qnzmax, int hmax, const T* shift, const T* post, const int* xmap,
                       const int* ymap, const T* x, int64_t ldx, T* y, int64_t ldy, int64_t n, int ncols, const T* dot_with,
                       T* dot_out, void* dot_ws, cudaStream_t st) {
  constexpr int R = kQpRows;
  constexpr int VEC = 16 / sizeof(T);
  constexpr int CW = 4 * VEC;
  MGP_CHECK_ARG(qwptr && qidx && qval && qrows && diag && hptr && hcol && x && y, "lap_spmm_qp: null pointer");
  MGP_CHECK_ARG(((uintptr_t)qwptr) % 16 == 0 && ((uintptr_t)hptr) % 16 == 0 && ((uintptr_t)hcol) % 16 == 0 && ((uintptr_t)qidx) % 16 == 0 &&
                    ((uintptr_t)qval) % 16 == 0,
                "lap_spmm_qp: qwptr / hptr / hcol / qidx / qval must be 16-byte aligned (bulk copies)");
  MGP_CHECK_ARG(hmax >= 0 && hmax % 4 == 0 && lmax >= tile_rows + hmax, "lap_spmm_qp: bad halo statistics hmax=%d lmax=%d", hmax, lmax);
  MGP_CHECK_ARG(tile_rows == R, "lap_spmm_qp: this build supports tile_rows == %d (got %d)", R, tile_rows);
  MGP_CHECK_ARG(n > 0 && ncols > 0 && ldx >= ncols && ldy >= ncols, "lap_spmm_qp: bad shape");
  MGP_CHECK_ARG(x != y, "lap_spmm_qp: X and Y must not alias");
  MGP_CHECK_ARG((dot_out == nullptr) || (dot_with && dot_ws), "lap_spmm_qp: dot epilogue needs dot_with and dot_ws");
  MGP_CHECK_ARG(lmax >= R && lmax <= 65535 && qnzmax >= 0 && qnzmax % 8 == 0, "lap_spmm_qp: bad tile statistics lmax=%d qnzmax=%d",
                lmax, qnzmax);
  const bool ok = (ncols % CW == 0) && (ldx % VEC == 0) && (ldy % VEC == 0) && (((uintptr_t)x) % 16 == 0) &&
                  (((uintptr_t)y) % 16 == 0) && (dot_with == nullptr || ((uintptr_t)dot_with) % 16 == 0);
  if (!ok) return MGP_EUNSUPPORTED;
  QpArgs<T> g;
  g.wptr = qwptr; g.qrows = qrows; g.wcol = qidx; g.aw = qval; g.diag = diag; g.hptr = hptr; g.hcol = hcol; g.shift = shift;
  g.post = post; g.xmap = xmap; g.ymap = ymap; g.x = x; g.ldx = ldx; g.y = y; g.ldy = ldy; g.n = n;
  g.peer_x = nullptr; g.npeers = 0; g.rank = 0; g.sync_flags = nullptr; g.sync_epoch = nullptr;
  g.ntiles = (int)ceil_div(n, (int64_t)R);
  g.lmax = (lmax + 3) & ~3;
  g.nzcap = qnzmax + 8;
  g.hmax = hmax > 0 ? hmax : 4;
  const size_t one = qp_stage_bytes<T>(g.lmax, g.nzcap);
  const size_t rings = qp_ring_bytes(g.hmax);
  g.stages = (3 * one + rings <= kQpSmemLimit) ? 3 : 2;
  const size_t smem = g.stages * one + rings;
  if (smem > kQpSmemLimit) return MGP_EUNSUPPORTED;
  g.dot_with = dot_out ? dot_with : nullptr;
  g.dot_out = dot_out;
  g.counter = dot_out ? reinterpret_cast<unsigned int*>(dot_ws) : nullptr;
  g.partials = dot_out ? reinterpret_cast<T*>(reinterpret_cast<char*>(dot_ws) + 256) : nullptr;
  g.dot_is_x = (dot_out && dot_with == x) ? 1 : 0;
  g.debug = 0;
  constexpr int PW = sizeof(T) == 4 ? 16 : 8;
  auto kern = lap_spmm_qp_kernel<T, PW>;
  static size_t configured = 0;   // per instantiation
  if (smem > configured) {
    MGP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  int64_t blocks = kNumSMs;
  if (blocks > g.ntiles) blocks = g.ntiles;
  for (int c0 = 0; c0 < ncols; c0 += CW) {
    g.c0 = c0;
    kern<<<(unsigned)blocks, (kQpConsumerWarps + PW) * 32, smem, st>>>(g);
    MGP_LAUNCH_CHECK();
  }
  return MGP_OK;
}

}  // namespace mgp

extern "C" {

int mgp_lap_spmm_qp_f32(const int32_t* qwptr, const uint16_t* qidx, const float* qval, const int32_t* qrows, const float* diag,
                        const int32_t* hptr, const int32_t* hcol, int32_t tile_rows, int32_t lmax, int32_t qnzmax, int32_t hmax,
                        const float* shift, const float* post, const int32_t* xmap, const int32_t* ymap, const float* x, int64_t ldx,
                        float* y, int64_t ldy, int64_t n, int32_t ncols, const float* dot_with, float* dot_out, void* dot_ws,
                        void* stream) {
  return mgp::lap_spmm_qp<float>(qwptr, qidx, qval, qrows, diag, hptr, hcol, tile_rows, lmax, qnzmax, hmax, shift, post, xmap, ymap, x,
                                 ldx, y, ldy, n, ncols, dot_with, dot_out, dot_ws, (cudaStream_t)stream);
}
int mgp_lap_spmm_qp_f64(const int32_t* qwptr, const uint16_t* qidx, const double* qval, const int32_t* qrows, const double* diag,
                        const int32_t* hptr, const int32_t* hcol, int32_t tile_rows, int32_t lmax, int32_t qnzmax, int32_t hmax,
                        const double* shift, const double* post, const int32_t* xmap, const int32_t* ymap, const double* x,
                        int64_t ldx, double* y, int64_t ldy, int64_t n, int32_t ncols, const double* dot_with, double* dot_out,
                        void* dot_ws, void* stream) {
  return mgp::lap_spmm_qp<double>(qwptr, qidx, qval, qrows, diag, hptr, hcol, tile_rows, lmax, qnzmax, hmax, shift, post, xmap, ymap, x,
                                  ldx, y, ldy, n, ncols, dot_with, dot_out, dot_ws, (cudaStream_t)stream);
}

}  // extern "C"
