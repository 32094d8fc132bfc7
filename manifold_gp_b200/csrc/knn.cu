// Exact brute-force kNN (squared L2, ascending), v1: fp32 CUDA-core distance tiles + fused threshold filter + rare-path
// warp-cooperative insertion.  Any d; bit-identical candidate ordering to the oracle's "direct" form
// (sum_d (q_d - x_d)^2 in ascending d, ties by ascending index).
//
// Reference: NearestNeighbors.search, manifold_gp/utils/nearest_neighbors.py:35-37 -> faiss Index{Flat,IVFFlat(nlist=1)}
// .search (exhaustive; riemann_kernel.py:40 builds the index with nlist=1).
//
// Shape: one CTA owns 64 query rows and sweeps the whole database in 64-point tiles.  256 threads each hold a 4x4
// micro-tile of squared distances in registers (d is consumed in chunks of 16 staged through shared memory).  Each
// distance is compared in-register with its row's current k-th best (tau); only survivors (~k ln(N/k) per row over the
// whole sweep) are pushed to a small shared-memory queue that the warps drain into per-row sorted lists.  The filter
// keeps the selection cost at ~1 compare per candidate, which is what bounds small-d searches (d = 3: the "contraction"
// is 3 FMAs per pair).  The tcgen05/TMEM variant for large d replaces only the distance-tile producer.
#include <float.h>

#include "common.cuh"

namespace mgp {

constexpr int kKnnQ = 64;        // query rows per CTA
constexpr int kKnnT = 64;        // database points per tile
constexpr int kKnnD = 16;        // dims per shared-memory chunk
constexpr int kKnnThreads = 256;
constexpr int kKnnQueue = 2048;  // survivor queue capacity per tile
constexpr int kKnnMaxK = 128;
constexpr int kKnnMaxParts = 32;  // database split of the list-mode re-search

struct KnnSmem {
  float qs[kKnnD][kKnnQ];
  float ds[kKnnD][kKnnT];
  float tau[kKnnQ];
  int qcount;
  float qd[kKnnQueue];
  int qi[kKnnQueue];
  int qr[kKnnQueue];
  // followed by: float listd[kKnnQ][k]; int listi[kKnnQ][k];
};

__device__ __forceinline__ bool lex_less(float d0, int i0, float d1, int i1) { return d0 < d1 || (d0 == d1 && i0 < i1); }

// Warp-cooperative insertion of (dnew, inew) into the ascending list of one row (k entries in shared memory).
__device__ __forceinline__ void knn_insert(float* ld, int* li, int k, float dnew, int inew, int lane, float* tau_r) {
  if (!lex_less(dnew, inew, ld[k - 1], li[k - 1])) return;  // warp-uniform (same smem values for every lane)
  int cnt = 0;
  for (int s = lane; s < k; s += 32) cnt += lex_less(ld[s], li[s], dnew, inew) ? 1 : 0;
  const int pos = warp_sum(cnt);
  float vd[kKnnMaxK / 32];
  int vi[kKnnMaxK / 32];
#pragma unroll
  for (int t = 0; t < kKnnMaxK / 32; ++t) {
    const int s = lane + 32 * t;
    if (s > pos && s < k) { vd[t] = ld[s - 1]; vi[t] = li[s - 1]; }
  }
  __syncwarp();
#pragma unroll
  for (int t = 0; t < kKnnMaxK / 32; ++t) {
    const int s = lane + 32 * t;
    if (s > pos && s < k) { ld[s] = vd[t]; li[s] = vi[t]; }
    if (s == pos) { ld[s] = dnew; li[s] = inew; }
  }
  __syncwarp();
  if (lane == 0) *tau_r = ld[k - 1];
}

// Drain `count` queued survivors into the row lists.  Warp w owns rows r with (r & 7) == w.
__device__ __forceinline__ void knn_drain(KnnSmem* s, float* listd, int* listi, int k, int count, int warp, int lane) {
  for (int base = 0; base < count; base += 32) {
    const int e = base + lane;
    const bool valid = e < count;
    const int row_e = valid ? s->qr[e] : -1;
    const float d_e = valid ? s->qd[e] : 0.f;
    const int i_e = valid ? s->qi[e] : 0;
    unsigned mask = __ballot_sync(0xffffffffu, valid && ((row_e & 7) == warp));
    while (mask) {
      const int src = __ffs(mask) - 1;
      mask &= mask - 1;
      const int r = __shfl_sync(0xffffffffu, row_e, src);
      const float dn = __shfl_sync(0xffffffffu, d_e, src);
      const int in = __shfl_sync(0xffffffffu, i_e, src);
      knn_insert(listd + r * k, listi + r * k, k, dn, in, lane, &s->tau[r]);
    }
  }
}

template <bool LIST>
__global__ void __launch_bounds__(kKnnThreads)
knn_kernel(const float* __restrict__ db, int64_t n, const float* __restrict__ q, int64_t nq, int d, int k,
           float* __restrict__ out_d, int64_t* __restrict__ out_i, const int* __restrict__ qlist,
           const unsigned int* __restrict__ qcount, float* __restrict__ part_d, int* __restrict__ part_i) {
  // qlist != NULL: the queries are qlist[0 .. *qcount) (the re-run of the queries the tensor-core search could not
  // certify, knn_tc.cu).  The launch has ceil(nq_max / 64) blocks; with nqb = ceil(count / 64) query blocks actually
  // needed, the spare blocks split the database Y = min(32, gridDim.x / nqb) ways: block b scans database part b / nqb for
  // query block b % nqb and writes its k best to part_d / part_i [Y][count][k]; knn_merge_parts_kernel merges them.
  int64_t blk = blockIdx.x;
  int ypart = 0, yparts = 1;
  if (LIST) {
    nq = (int64_t)*qcount;
    if (nq == 0) return;
    const int64_t nqb = (nq + kKnnQ - 1) / kKnnQ;
    yparts = (int)min((int64_t)kKnnMaxParts, (int64_t)gridDim.x / nqb);
    ypart = (int)(blk / nqb);
    if (ypart >= yparts) return;
    blk -= (int64_t)ypart * nqb;
  }
  extern __shared__ __align__(16) unsigned char smem_raw[];
  KnnSmem* s = reinterpret_cast<KnnSmem*>(smem_raw);
  float* listd = reinterpret_cast<float*>(smem_raw + sizeof(KnnSmem));
  int* listi = reinterpret_cast<int*>(listd + kKnnQ * k);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ty = tid >> 4, tx = tid & 15;  // 16 x 16 thread grid, 4 x 4 micro-tile each
  const int64_t q0 = blk * kKnnQ;

  for (int e = tid; e < kKnnQ * k; e += kKnnThreads) { listd[e] = FLT_MAX * 2.f; listi[e] = 0x7fffffff; }  // +inf
  if (tid < kKnnQ) s->tau[tid] = FLT_MAX * 2.f;
  if (tid == 0) s->qcount = 0;
  __syncthreads();

  const int64_t ntiles_all = (n + kKnnT - 1) / kKnnT;
  const int64_t tile_lo = LIST ? ntiles_all * ypart / yparts : 0;
  const int64_t ntiles = LIST ? ntiles_all * (ypart + 1) / yparts - tile_lo : ntiles_all;
  const int64_t t_first = LIST ? 0 : (q0 / kKnnT) % ntiles;  // start with the tile that holds the block's own rows (self-search)
  const bool single_chunk = d <= kKnnD;

  auto load_q_chunk = [&](int d0) {
    for (int e = tid; e < kKnnD * kKnnQ; e += kKnnThreads) {
      const int r = e % kKnnQ, dd = e / kKnnQ;
      const int64_t qe = q0 + r;
      const int64_t qi = qe < nq ? (LIST ? (int64_t)qlist[qe] : qe) : 0;
      s->qs[dd][r] = (qe < nq && d0 + dd < d) ? __ldg(q + qi * d + d0 + dd) : 0.f;
    }
  };
  if (single_chunk) load_q_chunk(0);

  for (int64_t tt = 0; tt < ntiles; ++tt) {
    const int64_t tile = tile_lo + (t_first + tt) % ntiles;
    const int64_t c0 = tile * kKnnT;
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;

    for (int d0 = 0; d0 < d; d0 += kKnnD) {
      __syncthreads();  // previous consumers of qs/ds (and the previous drain) are done
      if (!single_chunk) load_q_chunk(d0);
      for (int e = tid; e < kKnnD * kKnnT; e += kKnnThreads) {
        const int r = e % kKnnT, dd = e / kKnnT;
        const int64_t ci = c0 + r;
        s->ds[dd][r] = (ci < n && d0 + dd < d) ? __ldg(db + ci * d + d0 + dd) : 0.f;
      }
      __syncthreads();
      const int dmax = min(kKnnD, d - d0);
      for (int dd = 0; dd < dmax; ++dd) {
        const float4 qv = *reinterpret_cast<const float4*>(&s->qs[dd][ty * 4]);
        const float4 cv = *reinterpret_cast<const float4*>(&s->ds[dd][tx * 4]);
        const float qa[4] = {qv.x, qv.y, qv.z, qv.w};
        const float ca[4] = {cv.x, cv.y, cv.z, cv.w};
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            const float df = __fsub_rn(qa[a], ca[b]);
            acc[a][b] = __fadd_rn(acc[a][b], __fmul_rn(df, df));  // no FMA contraction: matches the oracle's mul-then-add
          }
      }
    }

    // ---- fused selection: in-register threshold filter, survivors to the queue --------------------------------
    float tau[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) tau[a] = s->tau[ty * 4 + a];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int64_t ci = c0 + tx * 4 + b;
        if (acc[a][b] <= tau[a] && ci < n && q0 + ty * 4 + a < nq) {
          const int pos = atomicAdd(&s->qcount, 1);
          if (pos < kKnnQueue) { s->qd[pos] = acc[a][b]; s->qi[pos] = (int)ci; s->qr[pos] = ty * 4 + a; }
        }
      }
    __syncthreads();
    const int count = s->qcount;
    if (count <= kKnnQueue) {
      if (count > 0) knn_drain(s, listd, listi, k, count, warp, lane);
      __syncthreads();
      if (tid == 0) s->qcount = 0;
    } else {
      // overflow (only while tau is still loose, i.e. the first tiles): redo the tile's selection in 16 sub-steps of
      // at most 256 survivors each; nothing from the overflowing attempt was inserted, so there are no duplicates.
      __syncthreads();
      if (tid == 0) s->qcount = 0;
      __syncthreads();
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const int64_t ci = c0 + tx * 4 + b;
          if (acc[a][b] <= s->tau[ty * 4 + a] && ci < n && q0 + ty * 4 + a < nq) {
            const int pos = atomicAdd(&s->qcount, 1);
            s->qd[pos] = acc[a][b]; s->qi[pos] = (int)ci; s->qr[pos] = ty * 4 + a;
          }
          __syncthreads();
          const int cnt2 = s->qcount;
          if (cnt2 > 0) knn_drain(s, listd, listi, k, cnt2, warp, lane);
          __syncthreads();
          if (tid == 0) s->qcount = 0;
          __syncthreads();
        }
    }
  }
  __syncthreads();
  for (int e = tid; e < kKnnQ * k; e += kKnnThreads) {
    const int r = e / k, c = e % k;
    const int64_t qe = q0 + r;
    if (qe < nq) {
      const int id = listi[e];
      const bool filled = id != 0x7fffffff;
      if (LIST) {
        part_d[((int64_t)ypart * nq + qe) * k + c] = filled ? listd[e] : __int_as_float(0x7f800000);
        part_i[((int64_t)ypart * nq + qe) * k + c] = id;   // 0x7fffffff = empty
      } else {
        out_d[qe * k + c] = filled ? listd[e] : __int_as_float(0x7f800000);
        out_i[qe * k + c] = filled ? (int64_t)id : (int64_t)-1;
      }
    }
  }
}

// One warp per re-searched query: merge its <= 32 sorted partial lists (one per lane) into the final k nearest.
__global__ void __launch_bounds__(128)
knn_merge_parts_kernel(const int* __restrict__ qlist, const unsigned int* __restrict__ qcount, int64_t nblocks_search, int k,
                       const float* __restrict__ part_d, const int* __restrict__ part_i, float* __restrict__ out_d,
                       int64_t* __restrict__ out_i) {
  const int64_t nq = (int64_t)*qcount;
  const int64_t e = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (e >= nq) return;
  const int64_t nqb = (nq + kKnnQ - 1) / kKnnQ;
  const int yparts = (int)min((int64_t)kKnnMaxParts, nblocks_search / nqb);
  const int64_t qi = qlist[e];
  const float kInf = __int_as_float(0x7f800000);
  const float* pd = part_d + ((int64_t)lane * nq + e) * k;
  const int* pi = part_i + ((int64_t)lane * nq + e) * k;
  int head = 0;
  float hd = lane < yparts ? pd[0] : kInf;
  int hi = lane < yparts ? pi[0] : 0x7fffffff;
  for (int c = 0; c < k; ++c) {
    float bd = hd;
    int bi = hi;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float od = __shfl_xor_sync(0xffffffffu, bd, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (lex_less(od, oi, bd, bi)) { bd = od; bi = oi; }
    }
    if (lane == 0) {
      out_d[qi * k + c] = bd;
      out_i[qi * k + c] = bi == 0x7fffffff ? (int64_t)-1 : (int64_t)bi;
    }
    if (bi != 0x7fffffff && hi == bi && hd == bd) {       // the winning lane advances (indices are unique across parts)
      ++head;
      hd = head < k ? pd[head] : kInf;
      hi = head < k ? pi[head] : 0x7fffffff;
    }
  }
}

}  // namespace mgp

namespace mgp {
// Exact search restricted to the query rows qlist[0 .. *qcount) (device memory); `nq_max` bounds the launch.
// part_d / part_i: scratch of knn_search_list_part_elems(nq_max, k) floats / ints.
int64_t knn_search_list_part_elems(int64_t nq_max, int k) { return ceil_div(nq_max, kKnnQ) * kKnnQ * k; }

int knn_search_list(const float* db, int64_t n, const float* q, int64_t nq_max, int d, int k, float* dist2, int64_t* idx,
                    const int* qlist, const unsigned int* qcount, float* part_d, int* part_i, cudaStream_t st) {
  const size_t smem = sizeof(KnnSmem) + (size_t)kKnnQ * k * 8;
  MGP_CUDA(cudaFuncSetAttribute(knn_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t grid = ceil_div(nq_max, kKnnQ);
  knn_kernel<true><<<(unsigned)grid, kKnnThreads, smem, st>>>(db, n, q, nq_max, d, k, dist2, idx, qlist, qcount, part_d, part_i);
  MGP_LAUNCH_CHECK();
  knn_merge_parts_kernel<<<(unsigned)ceil_div(nq_max, 4), 128, 0, st>>>(qlist, qcount, grid, k, part_d, part_i, dist2, idx);
  MGP_LAUNCH_CHECK();
  return MGP_OK;
}
}  // namespace mgp

using namespace mgp;

extern "C" {

size_t mgp_knn_search_ws_bytes(int64_t n, int64_t nq, int32_t d, int32_t k) {
  (void)n; (void)nq; (void)d; (void)k;
  return 256;  // v1 keeps all state in shared memory
}

int mgp_knn_search_f32(const float* db, int64_t n, const float* q, int64_t nq, int32_t d, int32_t k, float* dist2,
                       int64_t* idx, void* ws, size_t ws_bytes, void* stream) {
  (void)ws; (void)ws_bytes;
  MGP_CHECK_ARG(db && q && dist2 && idx, "knn_search: null pointer");
  MGP_CHECK_ARG(n > 0 && nq > 0 && d > 0, "knn_search: n, nq, d must be positive");
  MGP_CHECK_ARG(k > 0 && k <= kKnnMaxK, "knn_search: k must be in [1, %d] (got %d)", kKnnMaxK, k);
  MGP_CHECK_ARG(n < ((int64_t)1 << 31), "knn_search: n must be < 2^31");
  const size_t smem = sizeof(KnnSmem) + (size_t)kKnnQ * k * 8;
  MGP_CUDA(cudaFuncSetAttribute(knn_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t grid = ceil_div(nq, kKnnQ);
  MGP_CHECK_ARG(grid < ((int64_t)1 << 31), "knn_search: too many queries for one launch");
  knn_kernel<false><<<(unsigned)grid, kKnnThreads, smem, (cudaStream_t)stream>>>(db, n, q, nq, d, k, dist2, idx, nullptr, nullptr, nullptr, nullptr);
  MGP_LAUNCH_CHECK();
  return MGP_OK;
}

}  // extern "C"
