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
// Pipeline: producer warps fill a tile's region of a shared-memory BYTE RING -- bulk copies (TMA, UBLKCP) for the index stream,
// the value stream, the tile's own X rows, its slice of the diagonal and (paired walk) its row table, 16-byte cp.async for the
// scattered halo rows, all completing on the slot's mbarrier -- while 16 consumer warps walk the rows of a landed tile.
// Round 2, second half (all measured on B200, cfg-C, same-process A/B in profiles/):
//  * a tile occupies the bytes IT needs (mean 54 KB, worst 75 KB), handed out in ring order; up to 8 tiles in flight;
//  * the producers are specialised: ONE planner warp (metadata / halo-id rings, ring allocation, the slot's metadata block,
//    ahead of the fill), ONE filler warp (waits for the descriptor, the region and the ids, releases the helpers through a
//    named barrier, issues the bulk copies) and 14 helper warps that only issue the halo cp.async.  Before, all 16 producer
//    warps repeated the per-tile logic: 131 M warp instructions per launch, 26 M of them the walk; now 77 M.  Streaming with
//    the halo gather but without the walk went from 84 to 70 us, the kernel from 125 to 119 us (single-row walk);
//  * the paired-row walk (PAIR): 115 us.
#include "common.cuh"
#include "pipe_common.cuh"
#include "spmm_common.cuh"

namespace mgp {

constexpr int kWiSlots = 8;                        // tiles in flight at most (barrier pairs); their BYTES come from one ring, see below
constexpr int kWiSlotMeta = 24;                    // ints of per-slot metadata: [0,17) block offsets, 17 region offset, 18 / 19 value / index
                                                   // stream offsets (consumer-facing, written when the tile is filled); [20,24) the
                                                   // placement {region offset, X bytes, value bytes, tile to wait for} written two
                                                   // tiles ahead by the allocating thread
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
  unsigned int ring_bytes;   // bytes of the tile ring (multiple of 128)
  int hmax;      // max padded halo length (multiple of 4)
  const unsigned char* qrow;   // PAIR: tile-local output row of every (pair, half) position, [ntiles * 128]
  int c0;
  const T* dot_with;
  T* dot_out;
  T* partials;
  unsigned int* counter;
  int dot_is_x;
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
  int debug;     // timing experiments only (MGP_WI_DEBUG bit mask): 1 = consumers skip the row walk, 2 = producers skip the halo rows
                 // (same-process A/B on B200, cfg-C: 150 us full, 104 without halo copies, 90 without the walk, 62 with neither)
};

// Shared memory: [tile ring: ring_bytes][metadata ring: 2 chunks][halo-id ring: kWiIdSlots x hmax ints][slot metadata].
// A tile's region in the ring:  xs [(128 + nh) x 64 B] | vs [(cnt + 32) x VB] | cs [(cnt + 32) x u16]  rounded up to 128 B
// (nh = padded halo length, cnt = stream entries of the tile, VB = bytes of value(s) per entry; the 32 slack entries are what
// the consumers' look-ahead loads touch past the last warp block).  Regions are handed out in ring order by the producers
// -- a tile takes what IT needs, not the worst tile's footprint: at cfg-C the mean tile is 54 KB against a worst case of
// 75 KB, so ~3.5 tiles are in flight where fixed worst-case stages hold 2.
__host__ __device__ inline size_t wi_aux_bytes(int hmax) {
  return 2 * (size_t)kWiMetaBytes + (size_t)kWiIdSlots * hmax * 4 + (size_t)kWiSlots * kWiSlotMeta * 4 + (size_t)kWiSlots * 4;
}

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
//
// PAIR = true (fp32, "v6"): a slot is 8 lanes walking the UNION list of a row pair (graph.pair_streams): every 64-byte X-row
// load feeds both rows (32 FMAs per load instead of 16), entries carry two values (the lane's own output row, the other row).
// Lanes 0-3 of a slot end up with row A, lanes 4-7 with row B: one exchange across lane ^ 4 (16 shuffles), then the same
// 12-shuffle rotation as the single-row walk.  Morton-adjacent rows share a third of their columns: 0.73 stream slots and
// 0.43 shared-memory wavefronts per nonzero against 1.05 and 0.59.
template <typename T, int PW, bool PAIR>
__global__ void __launch_bounds__((kWiConsumerWarps + PW) * 32, 1)
lap_spmm_wi_kernel(const WiArgs<T> g) {
  static_assert(!PAIR || sizeof(T) == 4, "the paired walk is an fp32 kernel");
  constexpr int kWiProducerWarps = PW, kWiProducerThreads = PW * 32, kWiThreads = (kWiConsumerWarps + PW) * 32;
  constexpr int R = kWiRows;
  constexpr int VEC = 16 / sizeof(T);          // elements per 16-byte chunk
  constexpr int CW = 4 * VEC;                  // columns per pass: 64-byte rows
  constexpr uint32_t ROW_BYTES = 64;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr uint32_t VB = (PAIR ? 2u : 1u) * (uint32_t)sizeof(T);      // bytes of value(s) per stream entry
  // tail of a tile's region: the tile's slice of the diagonal [R] and (PAIR) its row table [R bytes] -- bulk-copied with the
  // streams so the consumers' epilogue has no global load on its critical path (ncu, round 2: 14 % of the paired walk's
  // consumer stall samples were the row-table load and the diagonal load behind it)
  constexpr uint32_t kTailBytes = (uint32_t)R * (uint32_t)sizeof(T) + (PAIR ? (uint32_t)R : 0u);
  __shared__ __align__(8) uint64_t full_bar[kWiSlots];
  __shared__ __align__(8) uint64_t empty_bar[kWiSlots];
  __shared__ __align__(8) uint64_t meta_bar[2];
  __shared__ __align__(8) uint64_t ids_bar[kWiIdSlots];
  __shared__ __align__(8) uint64_t desc_bar[kWiSlots];        // planner -> filler: the slot's metadata block is complete
  __shared__ __align__(8) uint64_t issued_bar[kWiIdSlots];    // helpers -> leader: all copies of the tile are issued (its id list is free)
  __shared__ const unsigned char* peer_tab[32];
  const uint32_t RB = g.ring_bytes;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  unsigned char* const ring = smem_raw + RB;                       // metadata ring
  int* const ids_ring = reinterpret_cast<int*>(ring + 2 * kWiMetaBytes);
  int* const slot_meta = ids_ring + (size_t)kWiIdSlots * g.hmax;

  pdl_wait();                                      // PDL: everything below may read what the predecessor launch wrote
  pdl_launch_dependents();                         // the successor's CTAs may become resident (they block in their own wait)
  if (g.done && *g.done != T(0)) return;          // uniform: CG converged, nothing to compute, publish or wait for
  if (g.peer_x && tid < g.npeers) peer_tab[tid] = g.peer_x[tid];
  if (tid == 0) {
    for (int s = 0; s < kWiSlots; ++s) {
      mbar_init(&full_bar[s], kWiProducerThreads - 64 + 1);   // the filler's expect_tx arrive + one cp.async arrive per helper thread
      mbar_init(&desc_bar[s], 1);
      mbar_init(&empty_bar[s], 16);                           // one arrival per warp block of the tile
    }
    for (int s = 0; s < kWiIdSlots; ++s) mbar_init(&issued_bar[s], kWiProducerThreads - 64);
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
    // =================================================== producers ==================================================
    // Three roles (see the header): planner (warp 1), filler (warp 0), helpers (warps 2 ..).  Why not every warp for itself:
    // the producer warps share the SM's issue slots with the consumers, and with every warp repeating ~300 instructions of
    // per-tile logic the producers issued more instructions than the row walk itself (+100 redundant instructions per tile
    // cost ~10 us per launch, measured with the first version of the ring allocator).
    const bool own_contig = (g.xmap == nullptr) && (g.ldx == CW);
    if (warp == 1) {
      // ------------------------------------- planner (two tiles ahead of the filler) --------------------------------
      // Everything that is NOT on the critical path of a tile's fill: the metadata / halo-id rings, the ring allocation and
      // the slot's metadata block.  One warp, so none of it is repeated; it posts desc_bar[slot] when the slot's block is
      // complete.  (Measured: with one thread doing this AND the fill, streaming alone took 82-84 us per launch against 65 us
      // for the round-1 kernel, whose fill thread did only waits + copies: the fill chain of one thread bounds the stream.)
      const int ch0 = t0 >> 5, ch_last = (t1 - 1) >> 5;             // metadata chunks this block touches
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
      // Ring allocation.  Regions are handed out at monotonically growing VIRTUAL addresses (physical = virtual mod RB; a
      // region never straddles the end of the ring: the remainder is skipped), so an older tile j must have been released
      // before tile i is written iff vlo_j < vhi_i - RB -- a condition monotone in j, tested with a running pointer.  The
      // previous user of the slot (tile i - 8) always counts.
      uint32_t a_phead = 0, a_vhead = 0;
      int a_tail = 0;
      int* const a_vlo = slot_meta + kWiSlots * kWiSlotMeta;          // [kWiSlots] virtual start of the last 8 tiles
      if (lane == 0 && t0 < t1) {
        request_meta(ch0);
        if (ch0 < ch_last) request_meta(ch0 + 1);
      }
      for (int t = t0; t < t1; ++t) {
        const int i = t - t0;
        const int s = i & (kWiSlots - 1);
        const int c = t >> 5;
        wait_meta(c);                                                // passes at once except on the first tile of a chunk
        const int* wp = meta_w(c) + 16 * (t & 31);
        const int* hp = meta_h(c) + (t & 31);
        const int base = wp[0];
        const uint32_t cnt = (uint32_t)(wp[16] - base), nh = (uint32_t)(hp[1] - hp[0]);
        const int rel = lane <= 16 ? wp[lane] - base : 0;
        const uint32_t xbytes = ((uint32_t)R + nh) * ROW_BYTES, vbytes = (cnt + 32u) * VB;
        const uint32_t size = (xbytes + vbytes + (cnt + 32u) * 2u + kTailBytes + 127u) & ~127u;
        // halo ids: the slot of tile i was last used by tile i - 4 -- free once every helper has issued that tile's copies
        if (lane == 0) {
          if (i >= kWiIdSlots) mbar_wait(&issued_bar[i % kWiIdSlots], (uint32_t)(((i / kWiIdSlots) - 1) & 1));
          request_ids(t);
        }
        // placement (all lanes compute the same values; lane 0's stores count)
        if (a_phead + size > RB) { a_vhead += RB - a_phead; a_phead = 0; }
        const uint32_t vhi = a_vhead + size;
        if (a_tail < i - (kWiSlots - 1)) a_tail = i - (kWiSlots - 1);
        while (a_tail < i && vhi > RB + (uint32_t)a_vlo[a_tail & (kWiSlots - 1)]) ++a_tail;
        __syncwarp();
        if (lane == 0) a_vlo[s] = (int)a_vhead;
        // the slot's metadata block was last read by the consumers of tile i - 8
        if (i >= kWiSlots) mbar_wait(&empty_bar[s], (uint32_t)((i / kWiSlots) - 1) & 1u);
        int* const sm = slot_meta + s * kWiSlotMeta;
        if (lane <= 16) sm[lane] = rel;
        else if (lane == 17) sm[17] = (int)a_phead;
        else if (lane == 18) sm[18] = (int)xbytes;
        else if (lane == 19) sm[19] = (int)(xbytes + vbytes);
        else if (lane == 20) sm[20] = base;
        else if (lane == 21) sm[21] = a_tail - 1;
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&desc_bar[s]);                                 // release: the filler may take the tile
          // (An L2 bulk prefetch of the streams 6 tiles ahead was measured: 113.3 vs 110.8 us without -- the fill is not bound by
          // the DRAM round trip.)
          if (((t + 1) & 31) == 0 && c + 2 <= ch_last) request_meta(c + 2);   // this chunk is finished: refill its ring slot
        }
        a_phead += size;
        a_vhead = vhi;
      }
    } else if (warp == 0) {
      // ------------------------------------------------ filler ------------------------------------------------
      // The critical path of the stream: wait for the tile's descriptor, its region (older tiles released) and its halo ids,
      // release the helpers, issue the bulk copies.
      int tail = 0;               // ordinal of the oldest tile whose release has not been waited for yet
      for (int t = t0; t < t1; ++t) {
        const int i = t - t0;
        const int s = i & (kWiSlots - 1);
        const int64_t row0 = (int64_t)t * R;
        const int nrows = (int)min((int64_t)R, g.n - row0);
        int* const sm = slot_meta + s * kWiSlotMeta;
        mbar_wait(&desc_bar[s], (uint32_t)(i / kWiSlots) & 1u);
        const int cnt = sm[16], need = sm[21];
        const uint32_t off = (uint32_t)sm[17], xbytes = (uint32_t)sm[18], cbytes0 = (uint32_t)sm[19];
        const int base = sm[20];
        // the region may be written once every older tile whose region it overlaps -- and the previous user of the slot --
        // has been released by all 16 consumer warps.  Tiles are released in order: the waits are a prefix of those in flight.
        for (; tail <= need; ++tail) mbar_wait(&empty_bar[tail & (kWiSlots - 1)], (uint32_t)(tail / kWiSlots) & 1u);
        mbar_wait(&ids_bar[i % kWiIdSlots], (uint32_t)((i / kWiIdSlots) & 1));
        unsigned char* const xs = smem_raw + off;
        unsigned char* const vs = xs + xbytes;
        unsigned char* const cs = xs + cbytes0;
        unsigned char* const dgs = cs + (size_t)(cnt + 32) * 2;       // diagonal slice, then the row table
        if (lane >= 28) {                                             // rows of a partial last tile past the 16-byte bulk copy
          const int e = (nrows & ~(int)(16 / sizeof(T) - 1)) + (lane - 28);
          if (e < nrows) reinterpret_cast<T*>(dgs)[e] = __ldg(g.diag + row0 + e);
        }
        // helpers: region, halo ids and metadata are ready.  A NAMED barrier, not an mbarrier: waiting helpers block in
        // hardware (ncu, round 2: 15 helper warps polling an mbarrier executed 42 M of the launch's instructions).  One barrier
        // id per slot: the filler is at most 8 tiles ahead of any helper (tile i + 8 needs tile i released, i.e. landed, i.e.
        // every helper past it), so an id never holds two pending arrivals of the filler.
        asm volatile("fence.acq_rel.cta;" ::: "memory");
        asm volatile("barrier.arrive %0, %1;" ::"r"(1 + s), "n"((PW - 1) * 32) : "memory");
        if (lane == 0) {
          const uint32_t dbytes = ((uint32_t)nrows * (uint32_t)sizeof(T)) & ~15u;
          const uint32_t bytes = (uint32_t)cnt * (2u + VB) + (own_contig ? (uint32_t)nrows * ROW_BYTES : 0u) + dbytes + (PAIR ? (uint32_t)R : 0u);
          mbar_arrive_expect_tx(&full_bar[s], bytes);
          if (cnt > 0) {
            bulk_g2s(vs, reinterpret_cast<const unsigned char*>(g.aw) + (size_t)base * VB, (uint32_t)cnt * VB, &full_bar[s]);
            bulk_g2s(cs, g.wcol + base, (uint32_t)cnt * 2, &full_bar[s]);
          }
          if (own_contig) bulk_g2s(xs, g.x + row0 * g.ldx + g.c0, (uint32_t)nrows * ROW_BYTES, &full_bar[s]);
          if (dbytes) bulk_g2s(dgs, g.diag + row0, dbytes, &full_bar[s]);
          if constexpr (PAIR) bulk_g2s(dgs + (size_t)R * sizeof(T), g.qrow + row0, (uint32_t)R, &full_bar[s]);
        }
        __syncwarp();
      }
    } else {
      // ------------------------------------------------ helpers -----------------------------------------------
      const int hw = warp - 2, sub = lane >> 2, ch = lane & 3;        // helper warp / row of a pass / 16-byte chunk
      const unsigned char* const xbase = reinterpret_cast<const unsigned char*>(g.x + g.c0) + ch * 16;
      const int64_t ldxb = g.ldx * (int64_t)sizeof(T);
      const int64_t xoff = (int64_t)g.c0 * (int64_t)sizeof(T) + ch * 16;   // same column window / chunk in a peer's X
      bool peers_ready = false;
      for (int t = t0; t < t1; ++t) {
        const int i = t - t0;
        const int s = i & (kWiSlots - 1);
        const int* const sm = slot_meta + s * kWiSlotMeta;
        const int64_t row0 = (int64_t)t * R;
        const int nown = own_contig ? 0 : (int)min((int64_t)R, g.n - row0);
        const int* ids = ids_ring + (size_t)(i % kWiIdSlots) * g.hmax;
        asm volatile("barrier.sync %0, %1;" ::"r"(1 + s), "n"((PW - 1) * 32) : "memory");
        unsigned char* const xs = smem_raw + (uint32_t)sm[17];
        const int nh = (sm[18] >> 6) - R;
        const int nscat = (g.debug & 2) ? 0 : nown + nh;
        // scattered X rows: 4 consecutive lanes copy the four 16-byte chunks of one row (8 whole rows per warp instruction).
        // Fused cross-GPU barrier: a warp waits for the peers' flags only when it reaches the first row that actually lives
        // on ANOTHER rank -- tiles away from the partition boundaries (most of them) never wait, so the NVLink round trip
        // of the barrier hides behind the interior tiles.
        for (int rb = 8 * hw; rb < nscat; rb += 8 * (kWiProducerWarps - 2)) {
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
        cp_async_arrive_noinc(&full_bar[s]);                          // arrives once this thread's copies have landed
        mbar_arrive(&issued_bar[i % kWiIdSlots]);                     // this thread is done with the tile's id list
      }
    }
  } else {
    // =================================================== consumers ==================================================
    const int rpos = lane >> 2;                      // output position inside the warp block: row slot (single-row walk) /
                                                     // 2 * pair slot + half (paired walk)
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
      const int64_t row0 = (int64_t)tile * R;
      int r = w * 8 + rpos;                          // row inside the tile (PAIR: position in the tile's row table, see below)
      const int s = ti & (kWiSlots - 1);
      const uint32_t ph = (uint32_t)(ti / kWiSlots) & 1u;
      const int* const rp = slot_meta + s * kWiSlotMeta;
      const int nrows = (int)min((int64_t)R, g.n - row0);
      // operands of the epilogue that live in global memory: in flight while the stage is awaited (single-row walk: the row
      // is known up front) and walked (paired walk: the row comes from the tile's row table, i.e. after the wait)
      bool active;
      int64_t row, yrow;
      Vec<T, VEC> dw;
      T po;
      auto issue_epilogue_loads = [&]() {
        active = r < nrows;
        row = row0 + r;
        if ((g.dot_out || g.ep_add) && !g.dot_is_x && active) {
          const int64_t drow = g.xmap ? (int64_t)__ldg(g.xmap + row) : row;
          dw = ldg_vec<T, VEC>(g.dot_with + drow * g.ldx + cbase);
        }
        yrow = row;
        if (g.ymap && active) yrow = (int64_t)__ldg(g.ymap + row);
        po = (g.post && active) ? __ldg(g.post + row) : T(1);
      };
      if constexpr (!PAIR) issue_epilogue_loads();
      mbar_wait(&full_bar[s], ph);                   // (sleep quanta of 32 / 96 / 320 ns between polls: no measurable difference)
      const int ofs = rp[w];
      const int steps = (g.debug & 1) ? 0 : (rp[w + 1] - ofs) >> 5;
      unsigned char* const xs = smem_raw + (uint32_t)rp[17];
      const unsigned short* cp = reinterpret_cast<const unsigned short*>(xs + (uint32_t)rp[19]) + ofs + lane;
      const unsigned char* const tailp = xs + (uint32_t)rp[19] + (uint32_t)(rp[16] + 32) * 2u;   // diagonal slice | row table
      if constexpr (PAIR) {
        r = (int)tailp[R * sizeof(T) + r];
        issue_epilogue_loads();
      }
      const T dgv = active ? reinterpret_cast<const T*>(tailp)[r] : T(0);
      T acc[4][VEC];
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int v = 0; v < VEC; ++v) acc[c][v] = T(0);
      uint32_t j = cp[0];
      if constexpr (PAIR) {
        const float2* vp = reinterpret_cast<const float2*>(xs + (uint32_t)rp[18]) + ofs + lane;
        float2 wv = vp[0];
        uint64_t am[4][2], ao[4][2];                 // partial sums of this lane's OWN output row / of the OTHER row of the pair
#pragma unroll
        for (int c = 0; c < 4; ++c) { am[c][0] = 0ull; am[c][1] = 0ull; ao[c][0] = 0ull; ao[c][1] = 0ull; }
        const uint32_t xb = smem_u32(xs);
        const uint32_t b0 = xb + o0, b1 = xb + o1, b2 = xb + o2, b3 = xb + o3;
#pragma unroll 2     // (4 steps in flight and a 32 / 96 register split were measured: no change, 115.0 vs 115.2 us)
        for (int t = 0; t < steps; ++t) {
          const uint32_t jn = cp[32];                // next step in flight (the last one reads the slack past the block: unused)
          const float2 wn = vp[32];
          cp += 32;
          vp += 32;
          const uint32_t off = j << 6;               // j * ROW_BYTES
          uint64_t x00, x01, x10, x11, x20, x21, x30, x31;
          lds_v2b64(b0 + off, x00, x01);
          lds_v2b64(b1 + off, x10, x11);
          lds_v2b64(b2 + off, x20, x21);
          lds_v2b64(b3 + off, x30, x31);
          const uint64_t wm = pack_f32x2(wv.x, wv.x), wo = pack_f32x2(wv.y, wv.y);
          ffma2_acc(am[0][0], wm, x00); ffma2_acc(am[0][1], wm, x01);
          ffma2_acc(am[1][0], wm, x10); ffma2_acc(am[1][1], wm, x11);
          ffma2_acc(am[2][0], wm, x20); ffma2_acc(am[2][1], wm, x21);
          ffma2_acc(am[3][0], wm, x30); ffma2_acc(am[3][1], wm, x31);
          ffma2_acc(ao[0][0], wo, x00); ffma2_acc(ao[0][1], wo, x01);
          ffma2_acc(ao[1][0], wo, x10); ffma2_acc(ao[1][1], wo, x11);
          ffma2_acc(ao[2][0], wo, x20); ffma2_acc(ao[2][1], wo, x21);
          ffma2_acc(ao[3][0], wo, x30); ffma2_acc(ao[3][1], wo, x31);
          j = jn;
          wv = wn;
        }
        // the other half of the slot (lane ^ 4) accumulated MY row in its ao: same chunk rotation (it shares l), so chunk c pairs up
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float m0, m1, m2, m3, q0, q1, q2, q3;
          unpack_f32x2(am[c][0], m0, m1); unpack_f32x2(am[c][1], m2, m3);
          unpack_f32x2(ao[c][0], q0, q1); unpack_f32x2(ao[c][1], q2, q3);
          acc[c][0] = m0 + __shfl_xor_sync(0xffffffffu, q0, 4);
          acc[c][1] = m1 + __shfl_xor_sync(0xffffffffu, q1, 4);
          acc[c][2] = m2 + __shfl_xor_sync(0xffffffffu, q2, 4);
          acc[c][3] = m3 + __shfl_xor_sync(0xffffffffu, q3, 4);
        }
      } else if constexpr (sizeof(T) == 4) {
        const T* vp = reinterpret_cast<const T*>(xs + (uint32_t)rp[18]) + ofs + lane;
        T wv = vp[0];
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
        const T* vp = reinterpret_cast<const T*>(xs + (uint32_t)rp[18]) + ofs + lane;
        T wv = vp[0];
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

// ---- value stream of the paired layout: out[i] = a[src[i]], 0 where src[i] < 0 (graph.pair_streams: two sources per entry) ----
template <typename T>
__global__ void __launch_bounds__(256)
lap_pair_values_kernel(const int* __restrict__ src, const T* __restrict__ a, int64_t count, T* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
    const int sidx = ld_stream(src + i);
    out[i] = sidx >= 0 ? __ldg(a + sidx) : T(0);
  }
}

template <typename T>
static int lap_pair_values(const int* src, const T* a, int64_t count, T* out, cudaStream_t st) {
  MGP_CHECK_ARG(src && a && out && count > 0, "lap_pair_values: bad arguments");
  const int64_t blocks = std::min<int64_t>(ceil_div(count, (int64_t)256), (int64_t)kNumSMs * 16);
  lap_pair_values_kernel<T><<<(unsigned)blocks, 256, 0, st>>>(src, a, count, out);
  MGP_LAUNCH_CHECK();
  return MGP_OK;
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
  g.hmax = hmax > 0 ? hmax : 4;
  g.qrow = nullptr;
  if (ext && ext->pair_rows) {
    if (sizeof(T) != 4) return MGP_EUNSUPPORTED;
    g.qrow = reinterpret_cast<const unsigned char*>(ext->pair_rows);
  }
  const bool pair = g.qrow != nullptr;
  const size_t entry_bytes = 2 + (pair ? 2 : 1) * sizeof(T);
  const size_t aux = wi_aux_bytes(g.hmax);
  // the ring takes all the shared memory there is: a tile occupies what it needs, so more bytes = more tiles in flight
  const size_t worst = ((size_t)(R + g.hmax) * 64 + (size_t)(wnzmax + 32) * entry_bytes + R * sizeof(T) + (pair ? R : 0) + 127) & ~(size_t)127;
  if (((uintptr_t)diag) % 16 != 0 || (pair && ((uintptr_t)g.qrow) % 16 != 0)) return MGP_EUNSUPPORTED;   // bulk copies of the tile tails
  if (aux + 128 > kWiSmemLimit) return MGP_EUNSUPPORTED;
  g.ring_bytes = (unsigned int)((kWiSmemLimit - aux) & ~(size_t)127);
  { const char* e = getenv("MGP_WI_RING_KB"); if (e && atoi(e) > 0) g.ring_bytes = std::min<unsigned int>(g.ring_bytes, (unsigned int)atoi(e) * 1024u); }
  if (worst > g.ring_bytes) return MGP_EUNSUPPORTED;
  const size_t smem = (size_t)g.ring_bytes + aux;
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
  auto kern = pw == 16 ? lap_spmm_wi_kernel<T, 16, false> : pw == 12 ? lap_spmm_wi_kernel<T, 12, false> : pw == 8 ? lap_spmm_wi_kernel<T, 8, false> : lap_spmm_wi_kernel<T, 4, false>;
  int kWiThreads = (kWiConsumerWarps + pw) * 32;
  if constexpr (sizeof(T) == 4) {
    if (pair) {
      kern = lap_spmm_wi_kernel<T, 16, true>;
      kWiThreads = (kWiConsumerWarps + 16) * 32;
    }
  }
  static size_t configured[2] = {0, 0};   // per instantiation (pw is fixed per process)
  if (smem > configured[pair ? 1 : 0]) {
    MGP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured[pair ? 1 : 0] = smem;
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

int mgp_lap_pair_values_f32(const int32_t* src, const float* a, int64_t count, float* out, void* stream) {
  return mgp::lap_pair_values<float>(src, a, count, out, (cudaStream_t)stream);
}
int mgp_lap_pair_values_f64(const int32_t* src, const double* a, int64_t count, double* out, void* stream) {
  return mgp::lap_pair_values<double>(src, a, count, out, (cudaStream_t)stream);
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
