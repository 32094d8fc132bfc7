// v5 SpMM ("warp-interleaved"): the pipelined tile kernel of lap_spmm_pipe.cu with the row work split the other way.
//
//   Y = post .* ( (diag + shift) .* X  -  A X )        for 64-byte rows of X (16 fp32 or 8 fp64 columns per pass)
//
// Why: ncu on v3/v4 (profiles/) shows the kernel bound by shared-memory wavefronts (LSU data pipe ~70 % busy): per 32
// nonzeros 16 wavefronts of X-row loads (64 B per nonzero -- irreducible) plus 6-8 wavefronts of (index, value) loads,
// because the 4 lanes that share a row all load the same index and value (a broadcast still costs one wavefront per
// 32-bit slice per phase).  Here the 4 lanes of a row slot take DIFFERENT nonzeros of the row and each accumulates all
// 16 columns; the slot's partial sums are combined once per row with 12 shuffles.  The (index, value) streams are stored
// in exactly the order the lanes consume them -- for warp w of a tile, step t, lane l: position wptr[16*tile + w] + 32 t + l
// holds nonzero 4t + (l & 3) of row 8w + (l >> 2) -- so every step is ONE 64-byte and ONE 128-byte fully coalesced
// shared load per warp: 2 wavefronts per 32 nonzeros instead of 6-8, and 29 instead of 36 (v4) / ~90 (v3) instructions.
//
// Bank conflicts of the X-row loads: lane l reads 16-byte chunk (c + l) mod 4 of its nonzero's row in load c, so the 4
// lanes of a slot cover 4 different chunks; the two slots of a quarter-warp stay in different 64-byte halves of the
// bank space as long as their rows' tile-local indices have opposite parity, which the entry order arranges (even rows
// list even indices first, odd rows odd indices first -- graph.py).  Padding entries (weight 0) carry an index of the
// parity their position expects.
//
// Pipeline (unchanged from v3): 4 producer warps fill one of up to 3 shared-memory stages per tile -- bulk copies (TMA,
// UBLKCP) for the index stream, the value stream and the tile's own X rows, 16-byte cp.async for the scattered halo rows,
// all completing on an mbarrier -- while 16 consumer warps walk the rows of a ready stage.
#include "common.cuh"
#include "pipe_common.cuh"
#include "spmm_common.cuh"

namespace mgp {

constexpr int kWiMaxStages = 3;
constexpr int kWiConsumerWarps = 16;
constexpr int kWiRows = 128;                       // rows per tile = 16 consumer warps x 8 row slots
constexpr int kWiIdSlots = 4;                      // halo-id ring: ids are requested this many tiles ahead
constexpr int kWiChunk = 32;                       // tiles per metadata chunk
constexpr int kWiMetaW = 16 * kWiChunk + 4;        // ints of wptr per chunk (513 used)
constexpr int kWiMetaH = kWiChunk + 4;             // ints of hptr per chunk (33 used)
constexpr int kWiMetaBytes = (kWiMetaW + kWiMetaH) * 4;
constexpr size_t kWiSmemLimit = 232448 - 10240;   // dynamic; static shared memory (dot epilogue, barriers) stays < 8 KB

template <typename T>
struct WiArgs {
  const int* wptr;               // [16 * ntiles + 1] stream offsets of the per-warp blocks (multiples of 32 entries); padded, see header
  const unsigned short* wcol;    // tile-local column per stream entry
  const T* aw;                   // value per stream entry (0 for padding)
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
  // ---- extended multi-GPU / solver hooks (mgp_lap_spmm_wi_ex; all optional) ----
  const T* done;                       // device scalar: the launch is a no-op when *done != 0 (CG past convergence)
  unsigned int* const* wait_flags;     // like sync_flags but WITHOUT the publish at kernel start: producers wait (lazily, at the
                                        // first remote halo row) until wait_flags[rank][src] >= epoch for every src
  unsigned int* const* publish_flags;  // after ALL rows of Y are written (last block to finish, last column pass):
                                        // publish_flags[dst][rank] = epoch for every dst ("my Y is complete")
  unsigned int* ticket;                // device counter for the completion ticket (zeroed once by the caller; self-resetting)
  T* const* red_ptrs;                  // with dot_out on the last column pass: the last block ships [dot_out | ship_extra] to slot
  unsigned int* const* red_flags;      //   [kind][epoch & 1][rank] of every rank's reduction buffer, then red_flags[dst][rank] = epoch
  const T* ship_extra;                 // second local sum to ship (kind 1), e.g. |r|^2 partials of the vector kernel; may be NULL
  int ship_ncols;                      // total number of columns to ship
  int last_pass;                       // 1 on the launch of the last column pass
  const T* ep_coef;                    // optional epilogue: Y <- coef * Y (+ ADD) with coef = *ep_coef; the dot epilogue sees the scaled Y
  int ep_add;                          // 1: ADD = dot_with (rows like X; no dot product on such a launch) -- y = add + coef * (A' x)
  int debug;     // timing experiments only (MGP_WI_DEBUG bit mask): 1 = consumers skip the row walk, 2 = producers skip the halo rows,
                 // 4 = one elected producer warp polls the barriers and releases the others through a named barrier (slower)
                 // (same-process A/B on B200, cfg-C: 150 us full, 104 without halo copies, 90 without the walk, 62 with neither)
};

template <typename T>
__host__ __device__ inline size_t wi_stage_bytes(int lmax, int nzcap) {
  // xs [lmax] x 64 B | vs [nzcap] T | cs [nzcap] u16 | rp [20] int       (every piece a multiple of 16 bytes)
  return (size_t)lmax * 64 + (size_t)nzcap * (sizeof(T) + 2) + 20 * 4;
}
__host__ __device__ inline size_t wi_ring_bytes(int hmax) { return 2 * (size_t)kWiMetaBytes + (size_t)kWiIdSlots * hmax * 4; }

// packed fp32 pairs (sm_100: fma.rn.f32x2 -> FFMA2)
__device__ __forceinline__ uint64_t pack_f32x2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ uint64_t pack_f32x2(double, double) { return 0ull; }     // never called (fp64 keeps scalar FMAs)
__device__ __forceinline__ void unpack_f32x2(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ void unpack_f32x2(uint64_t, double&, double&) {}
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}

__device__ __forceinline__ void ffma2_acc(uint64_t& c, uint64_t a, uint64_t b) {       // c += a * b (pairwise, in place)
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(c) : "l"(a), "l"(b));
}
__device__ __forceinline__ void lds_v2b64(uint32_t addr, uint64_t& lo, uint64_t& hi) {   // 16-byte shared load as two f32 pairs
  asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(lo), "=l"(hi) : "r"(addr));
}

template <int PW>
__device__ __forceinline__ void producers_sync() { asm volatile("bar.sync 1, %0;" ::"n"(PW * 32) : "memory"); }

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
__global__ void __launch_bounds__((kWiConsumerWarps + PW) * 32, 1)
lap_spmm_wi_kernel(const WiArgs<T> g) {
  constexpr int kWiProducerWarps = PW, kWiProducerThreads = PW * 32, kWiThreads = (kWiConsumerWarps + PW) * 32;
  constexpr int R = kWiRows;
  constexpr int VEC = 16 / sizeof(T);          // elements per 16-byte chunk
  constexpr int CW = 4 * VEC;                  // columns per pass: 64-byte rows
  constexpr uint32_t ROW_BYTES = 64;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kWiMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kWiMaxStages];
  __shared__ __align__(8) uint64_t meta_bar[2];
  __shared__ __align__(8) uint64_t ids_bar[kWiIdSlots];
  __shared__ const unsigned char* peer_tab[32];
  const int nstages = g.stages;
  const size_t stage_bytes = wi_stage_bytes<T>(g.lmax, g.nzcap);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  unsigned char* const ring = smem_raw + (size_t)nstages * stage_bytes;
  int* const ids_ring = reinterpret_cast<int*>(ring + 2 * kWiMetaBytes);

  pdl_wait();                                      // PDL: everything below may read what the predecessor launch wrote
  pdl_launch_dependents();                         // the successor's CTAs may become resident (they block in their own wait)
  if (g.done && *g.done != T(0)) return;          // uniform: CG converged, nothing to compute, publish or wait for
  if (g.peer_x && tid < g.npeers) peer_tab[tid] = g.peer_x[tid];
  if (tid == 0) {
    for (int s = 0; s < nstages; ++s) {
      mbar_init(&full_bar[s], 2 * kWiProducerThreads);   // per producer thread: one plain arrive (thread 0: expect_tx) + one cp.async arrive
      mbar_init(&empty_bar[s], 16);                      // one arrival per warp block of the tile
    }
    mbar_init(&meta_bar[0], 1);
    mbar_init(&meta_bar[1], 1);
    for (int s = 0; s < kWiIdSlots; ++s) mbar_init(&ids_bar[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  unsigned int sync_epoch = 0u;
  if (g.sync_epoch) sync_epoch = (unsigned int)(*g.sync_epoch) + 1u;
  if (g.sync_flags) {
    if (blockIdx.x == 0 && warp == 1 && lane < g.npeers) {      // "everything enqueued before this launch is done on my side"
      __threadfence_system();
      st_release_sys(g.sync_flags[lane] + g.rank, sync_epoch);
    }
  }
  unsigned int* const* const wait_tab = g.wait_flags ? g.wait_flags : g.sync_flags;

  // contiguous tile range of this block
  const int t0 = (int)(((int64_t)blockIdx.x * g.ntiles) / gridDim.x);
  const int t1 = (int)(((int64_t)(blockIdx.x + 1) * g.ntiles) / gridDim.x);

  T dsum[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) dsum[v] = T(0);

  // Register re-partition between the warp-specialised halves (PW == 16: 1024 threads = 64 registers each at launch): the
  // producers issue copies and need few registers, the consumers' walk is register-starved at 64 (ncu source page, round 2: the
  // unrolled step re-materialised the stage base and shuffled accumulator pairs with 8 IMAD.MOV per 2 steps).  setmaxnreg is
  // warpgroup-granular; producers = warpgroups 0..3, consumers = warpgroups 4..7; 16 x 32 x (40 + 88) = 65536 registers.
  if constexpr (PW == 16 && sizeof(T) == 4) {
    if (warp < kWiProducerWarps) asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    else asm volatile("setmaxnreg.inc.sync.aligned.u32 88;");
  }
  if (warp < kWiProducerWarps) {
    // =================================================== producer ===================================================
    const int pw = warp, sub = lane >> 2, ch = lane & 3;          // producer warp / row of a pass / 16-byte chunk
    const int ch0 = t0 >> 5, ch_last = (t1 - 1) >> 5;             // metadata chunks this block touches
    const bool own_contig = (g.xmap == nullptr) && (g.ldx == CW);
    const unsigned char* const xbase = reinterpret_cast<const unsigned char*>(g.x + g.c0) + ch * 16;
    const int64_t ldxb = g.ldx * (int64_t)sizeof(T);
    const int64_t xoff = (int64_t)g.c0 * (int64_t)sizeof(T) + ch * 16;   // same column window / chunk in a peer's X
    auto meta_w = [&](int c) { return reinterpret_cast<const int*>(ring + (size_t)((c - ch0) & 1) * kWiMetaBytes); };
    auto meta_h = [&](int c) { return meta_w(c) + kWiMetaW; };
    auto request_meta = [&](int c) {                               // one thread
      uint64_t* bar = &meta_bar[(c - ch0) & 1];
      mbar_arrive_expect_tx(bar, (uint32_t)kWiMetaBytes);
      bulk_g2s(const_cast<int*>(meta_w(c)), g.wptr + (size_t)c * 16 * kWiChunk, kWiMetaW * 4, bar);
      bulk_g2s(const_cast<int*>(meta_h(c)), g.hptr + (size_t)c * kWiChunk, kWiMetaH * 4, bar);
    };
    auto wait_meta = [&](int c) { mbar_wait(&meta_bar[(c - ch0) & 1], (uint32_t)(((c - ch0) >> 1) & 1)); };
    auto request_ids = [&](int t) {                                // one thread; the chunk of tile t must have landed
      const int i = t - t0;
      const int* hp = meta_h(t >> 5) + (t & 31);
      const int h0 = hp[0], nh = hp[1] - h0;
      uint64_t* bar = &ids_bar[i % kWiIdSlots];
      mbar_arrive_expect_tx(bar, (uint32_t)nh * 4u);
      if (nh > 0) bulk_g2s(ids_ring + (size_t)(i % kWiIdSlots) * g.hmax, g.hcol + h0, (uint32_t)nh * 4u, bar);
    };
    if (tid == 32 && t0 < t1) {                                    // prologue of the two rings
      request_meta(ch0);
      if (ch0 < ch_last) request_meta(ch0 + 1);
      wait_meta(ch0);
      for (int t = t0; t < t1 && t < t0 + kWiIdSlots; ++t) {
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
      unsigned short* const cs = reinterpret_cast<unsigned short*>(vs + g.nzcap);
      int* const rp = reinterpret_cast<int*>(cs + g.nzcap);
      const int c = t >> 5;
      // Experiment (MGP_WI_DEBUG bit 4): ONE producer warp polls for this tile's three conditions and releases the others
      // through the named barrier.  Measured in one process on B200 (cfg-C, C = 16): 124.9 us against 121.1 us with every
      // producer warp polling for itself -- the extra barrier per tile costs more than the polling it saves.  Off by default.
      if (g.debug & 4) {
        if (pw == 0) {
          wait_meta(c);
          mbar_wait(&ids_bar[i % kWiIdSlots], (uint32_t)((i / kWiIdSlots) & 1));
          mbar_wait(&empty_bar[s], ph ^ 1);
        }
        producers_sync<PW>();
      }
      wait_meta(c);                                                // passes at once except on the first tile of a chunk
      const int* wp = meta_w(c) + 16 * (t & 31);
      const int* hp = meta_h(c) + (t & 31);
      const int base = wp[0];
      const int cnt = wp[16] - base;                               // multiple of 32 entries
      const int nh = hp[1] - hp[0];
      const int64_t row0 = (int64_t)t * R;
      const int nrows = (int)min((int64_t)R, g.n - row0);
      const int nown = own_contig ? 0 : nrows;
      const int nscat = (g.debug & 2) ? 0 : nown + nh;
      const int* ids = ids_ring + (size_t)(i % kWiIdSlots) * g.hmax;
      mbar_wait(&ids_bar[i % kWiIdSlots], (uint32_t)((i / kWiIdSlots) & 1));
      mbar_wait(&empty_bar[s], ph ^ 1);                            // fresh barrier: parity 1 passes immediately
      if (tid == 0) {
        const uint32_t bytes = (uint32_t)cnt * (2 + (uint32_t)sizeof(T)) + (own_contig ? (uint32_t)nrows * ROW_BYTES : 0u);
        mbar_arrive_expect_tx(&full_bar[s], bytes);
        if (cnt > 0) {
          bulk_g2s(cs, g.wcol + base, (uint32_t)cnt * 2, &full_bar[s]);
          bulk_g2s(vs, g.aw + base, (uint32_t)cnt * (uint32_t)sizeof(T), &full_bar[s]);
        }
        if (own_contig) bulk_g2s(xs, g.x + row0 * g.ldx + g.c0, (uint32_t)nrows * ROW_BYTES, &full_bar[s]);
      }
      // scattered X rows: 4 consecutive lanes copy the four 16-byte chunks of one row (8 whole rows per warp instruction).
      // Fused cross-GPU barrier: a warp waits for the peers' flags only when it reaches the first row that actually lives
      // on ANOTHER rank -- tiles away from the partition boundaries (most of them) never wait, so the NVLink round trip
      // of the barrier hides behind the interior tiles.
      for (int rb = 8 * pw; rb < nscat; rb += 8 * kWiProducerWarps) {
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
        if (wait_tab && !peers_ready && __any_sync(0xffffffffu, remote)) {
          if (lane < g.npeers) {
            const unsigned int* f = wait_tab[g.rank] + lane;
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
      producers_sync<PW>();                                          // everyone is done with this tile's id slot and metadata
      if (tid == 32) {
        const int tn = t + kWiIdSlots;
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
    const int slot = lane >> 2;                      // row slot inside the warp
    const int l = lane & 3;                          // which nonzeros of the row (4t + l) / which output chunk
    const int cbase = g.c0 + l * VEC;
    const T shift = g.shift ? *g.shift : T(0);
    // byte offset of the chunk this lane reads in load c: chunk (c + l) mod 4
    const uint32_t o0 = (uint32_t)(((0 + l) & 3) * 16), o1 = (uint32_t)(((1 + l) & 3) * 16);
    const uint32_t o2 = (uint32_t)(((2 + l) & 3) * 16), o3 = (uint32_t)(((3 + l) & 3) * 16);
    // Fixed warp <-> block map: warp w walks rows 8w .. 8w+7 of every tile.  (Handing the 16 blocks of a tile out
    // dynamically -- shared-memory ticket counter -- was measured in the same process: 150.3 vs 149.6 us, no gain.)
    for (unsigned int tk = (unsigned int)(warp - kWiProducerWarps);; tk += 16) {
      const int ti = (int)(tk >> 4);
      if (ti >= t1 - t0) break;
      const int w = (int)(tk & 15u);                 // warp block of the tile: rows 8w .. 8w+7
      const int tile = t0 + ti;
      const int r = w * 8 + slot;                    // row inside the tile
      int s;
      uint32_t ph;
      if (nstages == 3) { s = ti % 3; ph = (uint32_t)(ti / 3) & 1u; } else { s = ti & 1; ph = (uint32_t)(ti >> 1) & 1u; }
      unsigned char* const sb = smem_raw + (size_t)s * stage_bytes;
      unsigned char* const xs = sb;
      const T* const vs = reinterpret_cast<const T*>(sb + (size_t)g.lmax * ROW_BYTES);
      const unsigned short* const cs = reinterpret_cast<const unsigned short*>(vs + g.nzcap);
      const int* const rp = reinterpret_cast<const int*>(cs + g.nzcap);
      const int64_t row0 = (int64_t)tile * R;
      const int nrows = (int)min((int64_t)R, g.n - row0);
      const bool active = r < nrows;
      const int64_t row = row0 + r;
      // operands of the epilogue that live in global memory: in flight while the stage is awaited and walked
      Vec<T, VEC> dw;
      if ((g.dot_out || g.ep_add) && !g.dot_is_x && active) {
        const int64_t drow = g.xmap ? (int64_t)__ldg(g.xmap + row) : row;
        dw = ldg_vec<T, VEC>(g.dot_with + drow * g.ldx + cbase);
      }
      int64_t yrow = row;
      if (g.ymap && active) yrow = (int64_t)__ldg(g.ymap + row);
      const T po = (g.post && active) ? __ldg(g.post + row) : T(1);
      const T dgv = active ? __ldg(g.diag + row) : T(0);
      mbar_wait(&full_bar[s], ph);
      const int ofs = rp[w];
      const int steps = (g.debug & 1) ? 0 : (rp[w + 1] - ofs) >> 5;
      const unsigned short* cp = cs + ofs + lane;
      const T* vp = vs + ofs + lane;
      T acc[4][VEC];
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int v = 0; v < VEC; ++v) acc[c][v] = T(0);
      uint32_t j = cp[0];
      T wv = vp[0];
      if constexpr (sizeof(T) == 4) {
        // fp32: packed FMAs (fma.rn.f32x2, SASS FFMA2) -- two IEEE fused multiply-adds per issue slot, bit-identical results.
        // The walk is issue-bound as much as shared-memory bound (ncu source page): 16 FFMA -> 8 FFMA2 per nonzero.
        uint64_t a2[4][2];
#pragma unroll
        for (int c = 0; c < 4; ++c) { a2[c][0] = 0ull; a2[c][1] = 0ull; }
        // 32-bit shared-window addresses, one base per chunk of this lane's rotation: a step is 1 shift + 4 adds of address
        // arithmetic (the generic-pointer form re-derived the stage base inside the loop)
        const uint32_t xb = smem_u32(xs);
        const uint32_t b0 = xb + o0, b1 = xb + o1, b2 = xb + o2, b3 = xb + o3;
#pragma unroll 2
        for (int t = 0; t < steps; ++t) {
          const uint32_t jn = cp[32];                // next step in flight (the last one reads into the next block: unused)
          const T wn = vp[32];
          cp += 32;
          vp += 32;
          const uint32_t off = j << 6;               // j * ROW_BYTES
          uint64_t x00, x01, x10, x11, x20, x21, x30, x31;
          lds_v2b64(b0 + off, x00, x01);
          lds_v2b64(b1 + off, x10, x11);
          lds_v2b64(b2 + off, x20, x21);
          lds_v2b64(b3 + off, x30, x31);
          const uint64_t w2 = pack_f32x2(wv, wv);
          ffma2_acc(a2[0][0], w2, x00); ffma2_acc(a2[0][1], w2, x01);
          ffma2_acc(a2[1][0], w2, x10); ffma2_acc(a2[1][1], w2, x11);
          ffma2_acc(a2[2][0], w2, x20); ffma2_acc(a2[2][1], w2, x21);
          ffma2_acc(a2[3][0], w2, x30); ffma2_acc(a2[3][1], w2, x31);
          j = jn;
          wv = wn;
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          unpack_f32x2(a2[c][0], acc[c][0], acc[c][1]);
          unpack_f32x2(a2[c][1], acc[c][2], acc[c][3]);
        }
      } else {
#pragma unroll 2
        for (int t = 0; t < steps; ++t) {
          const uint32_t jn = cp[32];                // next step in flight (the last one reads into the next block: unused)
          const T wn = vp[32];
          cp += 32;
          vp += 32;
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
          j = jn;
          wv = wn;
        }
      }
      // acc[c] of lane l holds chunk (c + l) mod 4 of the slot's partial sums; lane l ends up with chunk l complete:
      // its own acc[0] plus acc[4 - d] of the lane d places further (mod 4) in the slot, d = 1..3
      T res[VEC];
#pragma unroll
      for (int v = 0; v < VEC; ++v) res[v] = acc[0][v];
#pragma unroll
      for (int d = 1; d < 4; ++d) {
        const int src = (lane & ~3) | ((l + d) & 3);
#pragma unroll
        for (int v = 0; v < VEC; ++v) res[v] += __shfl_sync(0xffffffffu, acc[4 - d][v], src);
      }
      if (active) {
        const Vec<T, VEC> xi = *reinterpret_cast<const Vec<T, VEC>*>(xs + (size_t)r * ROW_BYTES + l * 16);
        const T d = dgv + shift;
        Vec<T, VEC> out;
#pragma unroll
        for (int v = 0; v < VEC; ++v) out.v[v] = po * (d * xi.v[v] - res[v]);
        if (g.ep_coef) {                                   // wrapper algebra folded into the launch: y = add + coef * y
          const T cf = __ldg(g.ep_coef);
          if (g.dot_is_x) dw = xi;
#pragma unroll
          for (int v = 0; v < VEC; ++v) out.v[v] = g.ep_add ? fma(cf, out.v[v], dw.v[v]) : cf * out.v[v];
        }
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
    const bool last = spmm_dot_epilogue<T, VEC, 4, CW, kWiThreads>(dsum, CW, g.c0, g.partials, g.counter, g.dot_out);
    // Fused all-reduce, first half: the block that completed the LOCAL sums ships them (and the caller's second sum) to every
    // rank's reduction buffer and raises this rank's flag there -- the consumer kernel only has to wait and add.
    if (last && g.red_ptrs && g.last_pass) {
      __syncthreads();                                      // dot_out of this pass is written (earlier passes: stream order)
      const int buf = (int)(sync_epoch & 1u);
      for (int i = tid; i < g.npeers * g.ship_ncols; i += kWiThreads) {
        const int dst = i / g.ship_ncols, c = i - dst * g.ship_ncols;
        T* base = g.red_ptrs[dst];
        base[(((size_t)0 * 2 + buf) * g.npeers + g.rank) * 128 + c] = g.dot_out[c];
        if (g.ship_extra) base[(((size_t)1 * 2 + buf) * g.npeers + g.rank) * 128 + c] = g.ship_extra[c];
      }
      __threadfence_system();
      __syncthreads();
      if (tid < g.npeers) st_release_sys(g.red_flags[tid] + g.rank, sync_epoch);
    }
  }
  if (g.publish_flags && g.last_pass) {
    // "Y complete" flag from the producing kernel itself (not from the first block of the consumer launch: that costs a
    // launch latency per sync point).  Every thread's stores are fenced, the last block to arrive publishes.
    __shared__ int publish_last;
    __threadfence();
    __syncthreads();
    if (tid == 0) {
      const unsigned int tk = atomicAdd(g.ticket, 1u);
      publish_last = (tk == gridDim.x - 1);
      if (publish_last) *g.ticket = 0u;
    }
    __syncthreads();
    if (publish_last && tid < g.npeers) {
      __threadfence_system();
      st_release_sys(g.publish_flags[tid] + g.rank, sync_epoch);
    }
  }
}

// ---- value stream in the warp-interleaved layout: aw[wptr[blk] + 32 t + lane] = a[rowptr[row] + 4t + (lane & 3)] or 0 ----
template <typename T>
__global__ void __launch_bounds__(256)
lap_wi_values_kernel(const int* __restrict__ rowptr, const int* __restrict__ wptr, const T* __restrict__ a, int64_t n,
                     int64_t nblocks, T* __restrict__ aw) {
  const int lane = threadIdx.x & 31;
  const int64_t blk = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (blk >= nblocks) return;
  const int64_t row = (blk >> 4) * kWiRows + (blk & 15) * 8 + (lane >> 2);
  int p0 = 0, len = 0;
  if (row < n) { p0 = rowptr[row]; len = rowptr[row + 1] - p0; }
  const int o = wptr[blk];
  const int steps = (wptr[blk + 1] - o) >> 5;
  for (int t = 0; t < steps; ++t) {
    const int e = 4 * t + (lane & 3);
    aw[(int64_t)o + 32 * t + lane] = e < len ? ld_stream(a + p0 + e) : T(0);
  }
}

template <typename T>
static int lap_wi_values(const int* rowptr, const int* wptr, const T* a, int64_t n, T* aw, cudaStream_t st) {
  MGP_CHECK_ARG(rowptr && wptr && a && aw && n > 0, "lap_wi_values: bad arguments");
  const int64_t nblocks = ceil_div(n, (int64_t)kWiRows) * 16;
  lap_wi_values_kernel<T><<<(unsigned)ceil_div(nblocks, (int64_t)8), 256, 0, st>>>(rowptr, wptr, a, n, nblocks, aw);
  MGP_LAUNCH_CHECK();
  return MGP_OK;
}

template <typename T>
static int lap_spmm_wi(const int* wptr, const unsigned short* wcol, const T* aw, const T* diag, const int* hptr,
                       const int* hcol, int tile_rows, int lmax, int wnzmax, int hmax, const T* shift, const T* post, const int* xmap,
                       const int* ymap, const T* x, int64_t ldx, T* y, int64_t ldy, int64_t n, int ncols, const T* dot_with,
                       T* dot_out, void* dot_ws, const void* peer_x, int npeers, int rank, const void* sync_flags, const T* sync_epoch,
                       cudaStream_t st, const mgp_wi_ext* ext = nullptr) {
  constexpr int R = kWiRows;
  constexpr int VEC = 16 / sizeof(T);
  constexpr int CW = 4 * VEC;
  MGP_CHECK_ARG(wptr && wcol && aw && diag && hptr && hcol && x && y, "lap_spmm_wi: null pointer");
  MGP_CHECK_ARG(((uintptr_t)wptr) % 16 == 0 && ((uintptr_t)hptr) % 16 == 0 && ((uintptr_t)hcol) % 16 == 0 && ((uintptr_t)wcol) % 16 == 0 &&
                    ((uintptr_t)aw) % 16 == 0,
                "lap_spmm_wi: wptr / hptr / hcol / wcol / aw must be 16-byte aligned (bulk copies)");
  MGP_CHECK_ARG(hmax >= 0 && hmax % 4 == 0 && lmax >= tile_rows + hmax, "lap_spmm_wi: bad halo statistics hmax=%d lmax=%d", hmax, lmax);
  MGP_CHECK_ARG(tile_rows == R, "lap_spmm_wi: this build supports tile_rows == %d (got %d)", R, tile_rows);
  MGP_CHECK_ARG(n > 0 && ncols > 0 && ldx >= ncols && ldy >= ncols, "lap_spmm_wi: bad shape");
  MGP_CHECK_ARG(x != y, "lap_spmm_wi: X and Y must not alias");
  MGP_CHECK_ARG((dot_out == nullptr) || (dot_with && dot_ws), "lap_spmm_wi: dot epilogue needs dot_with and dot_ws");
  MGP_CHECK_ARG(lmax >= R && lmax <= 65535 && wnzmax >= 0 && wnzmax % 32 == 0, "lap_spmm_wi: bad tile statistics lmax=%d wnzmax=%d",
                lmax, wnzmax);
  const bool ok = (ncols % CW == 0) && (ldx % VEC == 0) && (ldy % VEC == 0) && (((uintptr_t)x) % 16 == 0) &&
                  (((uintptr_t)y) % 16 == 0) && (dot_with == nullptr || ((uintptr_t)dot_with) % 16 == 0);
  if (!ok) return MGP_EUNSUPPORTED;
  WiArgs<T> g;
  g.wptr = wptr; g.wcol = wcol; g.aw = aw; g.diag = diag; g.hptr = hptr; g.hcol = hcol; g.shift = shift;
  g.post = post; g.xmap = xmap; g.ymap = ymap; g.x = x; g.ldx = ldx; g.y = y; g.ldy = ldy; g.n = n;
  g.peer_x = reinterpret_cast<const unsigned char* const*>(peer_x); g.npeers = npeers; g.rank = rank;
  g.sync_flags = reinterpret_cast<unsigned int* const*>(const_cast<void*>(sync_flags)); g.sync_epoch = sync_epoch;
  MGP_CHECK_ARG(sync_flags == nullptr || (peer_x && sync_epoch && rank >= 0 && rank < npeers), "lap_spmm_wi: fused barrier needs peer X, an epoch pointer and a valid rank");
  MGP_CHECK_ARG(peer_x == nullptr || (npeers >= 1 && npeers <= 32 && xmap == nullptr), "lap_spmm_wi: peer X needs 1..32 ranks and no xmap");
  g.done = nullptr; g.wait_flags = nullptr; g.publish_flags = nullptr; g.ticket = nullptr; g.red_ptrs = nullptr;
  g.red_flags = nullptr; g.ship_extra = nullptr; g.ship_ncols = 0; g.last_pass = 0; g.ep_coef = nullptr; g.ep_add = 0;
  if (ext) {
    g.done = reinterpret_cast<const T*>(ext->done_flag);
    g.wait_flags = reinterpret_cast<unsigned int* const*>(const_cast<void*>(ext->wait_flags));
    g.publish_flags = reinterpret_cast<unsigned int* const*>(const_cast<void*>(ext->publish_flags));
    g.ticket = reinterpret_cast<unsigned int*>(ext->ticket);
    g.red_ptrs = reinterpret_cast<T* const*>(const_cast<void*>(ext->red_ptrs));
    g.red_flags = reinterpret_cast<unsigned int* const*>(const_cast<void*>(ext->red_flags));
    g.ship_extra = reinterpret_cast<const T*>(ext->ship_extra);
    g.ship_ncols = ext->ship_ncols;
    g.ep_coef = reinterpret_cast<const T*>(ext->ep_coef);
    g.ep_add = ext->ep_add ? 1 : 0;
    MGP_CHECK_ARG(!g.ep_add || (g.ep_coef && dot_with && !dot_out), "lap_spmm_wi_ex: ep_add needs ep_coef and the ADD operand in dot_with, and excludes the dot epilogue");
    const bool sync_any = g.wait_flags || g.publish_flags || g.red_ptrs;
    MGP_CHECK_ARG(!sync_any || (peer_x && sync_epoch && rank >= 0 && rank < npeers), "lap_spmm_wi_ex: flags need peer X, an epoch pointer and a valid rank");
    MGP_CHECK_ARG(g.publish_flags == nullptr || g.ticket != nullptr, "lap_spmm_wi_ex: publish_flags needs a ticket counter");
    MGP_CHECK_ARG(g.red_ptrs == nullptr || (g.red_flags && dot_out && g.ship_ncols > 0 && g.ship_ncols <= 128 && g.ship_ncols <= ncols),
                  "lap_spmm_wi_ex: shipping the dot partials needs red_flags, dot_out and 0 < ship_ncols <= min(128, ncols)");
    MGP_CHECK_ARG(sync_flags == nullptr || g.wait_flags == nullptr, "lap_spmm_wi_ex: sync_flags and wait_flags are exclusive");
    if (ext->publish_at_start && g.wait_flags) {      // the round-1 barrier form: block 0 publishes at kernel start, lazy wait
      g.sync_flags = g.wait_flags;
      g.wait_flags = nullptr;
    }
  }
  g.ntiles = (int)ceil_div(n, (int64_t)R);
  g.lmax = (lmax + 3) & ~3;
  g.nzcap = wnzmax + 32;                       // the consumers' look-ahead loads read one step past a warp block
  g.hmax = hmax > 0 ? hmax : 4;
  const size_t one = wi_stage_bytes<T>(g.lmax, g.nzcap);
  const size_t rings = wi_ring_bytes(g.hmax);
  g.stages = (3 * one + rings <= kWiSmemLimit) ? 3 : 2;
  const size_t smem = g.stages * one + rings;
  if (smem > kWiSmemLimit) return MGP_EUNSUPPORTED;
  g.dot_with = (dot_out || g.ep_add) ? dot_with : nullptr;
  g.dot_out = dot_out;
  g.counter = dot_out ? reinterpret_cast<unsigned int*>(dot_ws) : nullptr;
  g.partials = dot_out ? reinterpret_cast<T*>(reinterpret_cast<char*>(dot_ws) + 256) : nullptr;
  g.dot_is_x = ((dot_out || g.ep_add) && dot_with == x) ? 1 : 0;
  { const char* dbg = getenv("MGP_WI_DEBUG"); g.debug = dbg ? atoi(dbg) : 0; }
  // producer warps: measured on B200 (cfg-C, C = 16, fp32): 4 -> 163.7 us, 8 -> 144.0, 12 -> 133.4, 16 -> 127.0 (the halo
  // cp.async issue shares the LSU queue with the consumers' shared-memory loads; more producer warps = a larger share)
  static const int pw = [] {
    const char* e = getenv("MGP_WI_PW");
    const int v = e ? atoi(e) : (sizeof(T) == 4 ? 16 : 8);
    if (v == 16) return sizeof(T) == 4 ? 16 : 12;      // fp64 at 1024 threads would spill (64 registers)
    return (v == 4 || v == 8 || v == 12) ? v : 8;
  }();
  auto kern = pw == 16 ? lap_spmm_wi_kernel<T, 16> : pw == 12 ? lap_spmm_wi_kernel<T, 12> : pw == 8 ? lap_spmm_wi_kernel<T, 8> : lap_spmm_wi_kernel<T, 4>;
  const int kWiThreads = (kWiConsumerWarps + pw) * 32;
  static size_t configured = 0;   // per instantiation
  if (smem > configured) {
    MGP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  int64_t blocks = kNumSMs;
  if (blocks > g.ntiles) blocks = g.ntiles;
  for (int c0 = 0; c0 < ncols; c0 += CW) {
    g.c0 = c0;
    g.last_pass = (c0 + CW >= ncols) ? 1 : 0;
    MGP_CUDA(launch_pdl(kern, dim3((unsigned)blocks), dim3(kWiThreads), smem, st, g));
    MGP_LAUNCH_CHECK();
  }
  return MGP_OK;
}

}  // namespace mgp

extern "C" {

int mgp_lap_wi_values_f32(const int32_t* rowptr, const int32_t* wptr, const float* a, int64_t n, float* aw, void* stream) {
  return mgp::lap_wi_values<float>(rowptr, wptr, a, n, aw, (cudaStream_t)stream);
}
int mgp_lap_wi_values_f64(const int32_t* rowptr, const int32_t* wptr, const double* a, int64_t n, double* aw, void* stream) {
  return mgp::lap_wi_values<double>(rowptr, wptr, a, n, aw, (cudaStream_t)stream);
}

int mgp_lap_spmm_wi_f32(const int32_t* wptr, const uint16_t* wcol, const float* aw, const float* diag, const int32_t* hptr,
                        const int32_t* hcol, int32_t tile_rows, int32_t lmax, int32_t wnzmax, int32_t hmax, const float* shift,
                        const float* post, const int32_t* xmap, const int32_t* ymap, const float* x, int64_t ldx, float* y,
                        int64_t ldy, int64_t n, int32_t ncols, const float* dot_with, float* dot_out, void* dot_ws,
                        const void* peer_x, int32_t npeers, int32_t rank, const void* sync_flags, const float* sync_epoch,
                        void* stream) {
  return mgp::lap_spmm_wi<float>(wptr, wcol, aw, diag, hptr, hcol, tile_rows, lmax, wnzmax, hmax, shift, post, xmap, ymap, x,
                                 ldx, y, ldy, n, ncols, dot_with, dot_out, dot_ws, peer_x, npeers, rank, sync_flags, sync_epoch, (cudaStream_t)stream);
}
int mgp_lap_spmm_wi_f64(const int32_t* wptr, const uint16_t* wcol, const double* aw, const double* diag,
                        const int32_t* hptr, const int32_t* hcol, int32_t tile_rows, int32_t lmax, int32_t wnzmax, int32_t hmax,
                        const double* shift, const double* post, const int32_t* xmap, const int32_t* ymap, const double* x,
                        int64_t ldx, double* y, int64_t ldy, int64_t n, int32_t ncols, const double* dot_with, double* dot_out,
                        void* dot_ws, const void* peer_x, int32_t npeers, int32_t rank, const void* sync_flags, const double* sync_epoch,
                        void* stream) {
  return mgp::lap_spmm_wi<double>(wptr, wcol, aw, diag, hptr, hcol, tile_rows, lmax, wnzmax, hmax, shift, post, xmap, ymap, x,
                                  ldx, y, ldy, n, ncols, dot_with, dot_out, dot_ws, peer_x, npeers, rank, sync_flags, sync_epoch, (cudaStream_t)stream);
}
int mgp_lap_spmm_wi_ex_f32(const int32_t* wptr, const uint16_t* wcol, const float* aw, const float* diag, const int32_t* hptr,
                           const int32_t* hcol, int32_t tile_rows, int32_t lmax, int32_t wnzmax, int32_t hmax, const float* shift,
                           const float* post, const int32_t* xmap, const int32_t* ymap, const float* x, int64_t ldx, float* y,
                           int64_t ldy, int64_t n, int32_t ncols, const float* dot_with, float* dot_out, void* dot_ws,
                           const void* peer_x, int32_t npeers, int32_t rank, const float* sync_epoch, const mgp_wi_ext* ext,
                           void* stream) {
  return mgp::lap_spmm_wi<float>(wptr, wcol, aw, diag, hptr, hcol, tile_rows, lmax, wnzmax, hmax, shift, post, xmap, ymap, x,
                                 ldx, y, ldy, n, ncols, dot_with, dot_out, dot_ws, peer_x, npeers, rank, nullptr, sync_epoch, (cudaStream_t)stream, ext);
}
int mgp_lap_spmm_wi_ex_f64(const int32_t* wptr, const uint16_t* wcol, const double* aw, const double* diag, const int32_t* hptr,
                           const int32_t* hcol, int32_t tile_rows, int32_t lmax, int32_t wnzmax, int32_t hmax, const double* shift,
                           const double* post, const int32_t* xmap, const int32_t* ymap, const double* x, int64_t ldx, double* y,
                           int64_t ldy, int64_t n, int32_t ncols, const double* dot_with, double* dot_out, void* dot_ws,
                           const void* peer_x, int32_t npeers, int32_t rank, const double* sync_epoch, const mgp_wi_ext* ext,
                           void* stream) {
  return mgp::lap_spmm_wi<double>(wptr, wcol, aw, diag, hptr, hcol, tile_rows, lmax, wnzmax, hmax, shift, post, xmap, ymap, x,
                                  ldx, y, ldy, n, ncols, dot_with, dot_out, dot_ws, peer_x, npeers, rank, nullptr, sync_epoch, (cudaStream_t)stream, ext);
}

}  // extern "C"
