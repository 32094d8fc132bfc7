// Graph construction on the device: symmetrise + mean-coalesce the directed kNN lists, build the row-major
// directed structure (CSR over both directions) the SpMM kernels stream.
//
// Reference behaviour reproduced: manifold_gp/utils/nearest_neighbors.py:39-55 (drop self column, (min,max) map,
// torch_sparse.coalesce(op='mean') -> lexicographically sorted unique upper-triangular COO).
// Sorting uses CUB's device radix sort (a utility, not a hot op: once per graph, hyper-parameter independent).
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "common.cuh"

namespace mgp {

static inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

static inline int key_bits(int64_t n) {
  // keys are row*n + col < n*n (plus the all-ones sentinel handled separately)
  int b = 1;
  unsigned __int128 lim = (unsigned __int128)n * (unsigned __int128)n;
  while (b < 64 && (((unsigned __int128)1) << b) < lim) ++b;
  return b;
}

constexpr unsigned long long kSentinel = ~0ull;

// ---- symmetrise ------------------------------------------------------------------------------------------------
__global__ void sym_make_keys_kernel(const float* __restrict__ dist2, const int64_t* __restrict__ idx, int64_t n, int k,
                                     int drop, unsigned long long* __restrict__ keys, float* __restrict__ vals) {
  const int kk = k - drop;
  const int64_t total = n * kk;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = t / kk;
    const int c = (int)(t - i * kk) + drop;
    const int64_t j = idx[i * k + c];
    unsigned long long key = kSentinel;
    if (j >= 0 && j < n) {
      const int64_t lo = j > i ? i : j;  // split = cols > rows (nearest_neighbors.py:48-50)
      const int64_t hi = j > i ? j : i;
      key = (unsigned long long)lo * (unsigned long long)n + (unsigned long long)hi;
    }
    keys[t] = key;
    vals[t] = dist2[i * k + c];
  }
}

__global__ void sym_flag_heads_kernel(const unsigned long long* __restrict__ keys, int64_t total, int* __restrict__ flags) {
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const unsigned long long kcur = keys[t];
    flags[t] = (kcur != kSentinel && (t == 0 || keys[t - 1] != kcur)) ? 1 : 0;
  }
}

__global__ void sym_emit_kernel(const unsigned long long* __restrict__ keys, const float* __restrict__ vals,
                                const int* __restrict__ flags, const int* __restrict__ pos, int64_t total, int64_t n,
                                int64_t cap, int64_t* __restrict__ eidx, float* __restrict__ eval,
                                int64_t* __restrict__ m_out) {
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    if (flags[t]) {
      const unsigned long long kcur = keys[t];
      float s = vals[t];
      int cnt = 1;
      for (int64_t u = t + 1; u < total && keys[u] == kcur; ++u) {  // segments are 1-2 long (directed copies of one edge)
        s += vals[u];
        ++cnt;
      }
      const int64_t o = pos[t];
      eidx[o] = (int64_t)(kcur / (unsigned long long)n);
      eidx[cap + o] = (int64_t)(kcur % (unsigned long long)n);
      eval[o] = s / (float)cnt;  // op='mean'
    }
    if (t == total - 1) *m_out = (int64_t)pos[t] + (int64_t)flags[t];
  }
}

struct SymWs {
  unsigned long long *keys_a, *keys_b;
  float *vals_a, *vals_b;
  int *flags, *pos;
  void* cub;
  size_t cub_bytes;
  size_t total_bytes;
};

static int sym_layout(int64_t total, void* base, SymWs* w) {
  size_t sort_b = 0, scan_b = 0;
  cudaError_t e = cub::DeviceRadixSort::SortPairs(nullptr, sort_b, (unsigned long long*)nullptr, (unsigned long long*)nullptr,
                                                  (float*)nullptr, (float*)nullptr, total, 0, 64, (cudaStream_t)0);
  if (e != cudaSuccess) { set_error("cub sort size query: %s", cudaGetErrorString(e)); return MGP_ECUDA; }
  e = cub::DeviceScan::ExclusiveSum(nullptr, scan_b, (int*)nullptr, (int*)nullptr, total, (cudaStream_t)0);
  if (e != cudaSuccess) { set_error("cub scan size query: %s", cudaGetErrorString(e)); return MGP_ECUDA; }
  size_t off = 0;
  char* b = (char*)base;
  auto take = [&](size_t bytes) { char* p = b ? b + off : nullptr; off += align_up(bytes); return (void*)p; };
  w->keys_a = (unsigned long long*)take(total * 8);
  w->keys_b = (unsigned long long*)take(total * 8);
  w->vals_a = (float*)take(total * 4);
  w->vals_b = (float*)take(total * 4);
  w->flags = (int*)take(total * 4);
  w->pos = (int*)take(total * 4);
  w->cub_bytes = sort_b > scan_b ? sort_b : scan_b;
  w->cub = take(w->cub_bytes);
  w->total_bytes = off;
  return MGP_OK;
}

// ---- CSR build -----------------------------------------------------------------------------------------------
__global__ void csr_make_keys_kernel(const int64_t* __restrict__ eidx, int64_t ld, int64_t m, int64_t n,
                                     unsigned long long* __restrict__ keys, int* __restrict__ vals) {
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < m; e += (int64_t)gridDim.x * blockDim.x) {
    const unsigned long long i = (unsigned long long)eidx[e];
    const unsigned long long j = (unsigned long long)eidx[ld + e];
    keys[2 * e] = i * (unsigned long long)n + j;
    keys[2 * e + 1] = j * (unsigned long long)n + i;
    vals[2 * e] = (int)e;
    vals[2 * e + 1] = (int)e;
  }
}

__global__ void csr_emit_kernel(const unsigned long long* __restrict__ keys, int64_t nnz, int64_t n,
                                int* __restrict__ col) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < nnz; p += (int64_t)gridDim.x * blockDim.x)
    col[p] = (int)(keys[p] % (unsigned long long)n);
}

__global__ void csr_rowptr_kernel(const unsigned long long* __restrict__ keys, int64_t nnz, int64_t n,
                                  int* __restrict__ rowptr) {
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r <= n; r += (int64_t)gridDim.x * blockDim.x) {
    const unsigned long long target = (unsigned long long)r * (unsigned long long)n;  // first key of row r
    int64_t lo = 0, hi = nnz;
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if (keys[mid] < target) lo = mid + 1; else hi = mid;
    }
    rowptr[r] = (int)lo;
  }
}

struct CsrWs {
  unsigned long long *keys_a, *keys_b;
  int* vals_a;
  void* cub;
  size_t cub_bytes;
  size_t total_bytes;
};

static int csr_layout(int64_t nnz, void* base, CsrWs* w) {
  size_t sort_b = 0;
  cudaError_t e = cub::DeviceRadixSort::SortPairs(nullptr, sort_b, (unsigned long long*)nullptr, (unsigned long long*)nullptr,
                                                  (int*)nullptr, (int*)nullptr, nnz, 0, 64, (cudaStream_t)0);
  if (e != cudaSuccess) { set_error("cub sort size query: %s", cudaGetErrorString(e)); return MGP_ECUDA; }
  size_t off = 0;
  char* b = (char*)base;
  auto take = [&](size_t bytes) { char* p = b ? b + off : nullptr; off += align_up(bytes); return (void*)p; };
  w->keys_a = (unsigned long long*)take(nnz * 8);
  w->keys_b = (unsigned long long*)take(nnz * 8);
  w->vals_a = (int*)take(nnz * 4);
  w->cub_bytes = sort_b;
  w->cub = take(sort_b);
  w->total_bytes = off;
  return MGP_OK;
}

template <typename T>
__global__ void gather_edge_kernel(const T* __restrict__ val, const int* __restrict__ eid, int64_t nnz, T* __restrict__ out) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < nnz; p += (int64_t)gridDim.x * blockDim.x)
    out[p] = val[eid[p]];
}

static inline int grid_for(int64_t work, int block = 256) {
  int64_t g = ceil_div(work, block);
  const int64_t cap = (int64_t)kNumSMs * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace mgp

using namespace mgp;

extern "C" {

size_t mgp_graph_symmetrize_ws_bytes(int64_t n, int32_t k) {
  SymWs w;
  if (n <= 0 || k <= 0) return 0;
  if (sym_layout(n * (int64_t)k, nullptr, &w) != MGP_OK) return 0;
  return w.total_bytes;
}

int mgp_graph_symmetrize_f32(const float* dist2, const int64_t* idx, int64_t n, int32_t k, int32_t drop_first,
                             int64_t* eidx, float* eval, int64_t cap, int64_t* m_out, void* ws, size_t ws_bytes,
                             void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int drop = drop_first ? 1 : 0;
  MGP_CHECK_ARG(dist2 && idx && eidx && eval && m_out && ws, "graph_symmetrize: null pointer");
  MGP_CHECK_ARG(n > 0 && k > drop, "graph_symmetrize: need n > 0 and k > drop_first (n=%lld k=%d)", (long long)n, k);
  const int64_t total = n * (int64_t)(k - drop);
  MGP_CHECK_ARG(cap >= total, "graph_symmetrize: cap %lld < n*(k-drop) %lld", (long long)cap, (long long)total);
  MGP_CHECK_ARG(total < (int64_t)1 << 31, "graph_symmetrize: n*(k-1) must be < 2^31");
  SymWs w;
  int rc = sym_layout(total, ws, &w);
  if (rc != MGP_OK) return rc;
  if (ws_bytes < w.total_bytes) { set_error("graph_symmetrize: workspace %zu < %zu", ws_bytes, w.total_bytes); return MGP_EWORKSPACE; }

  sym_make_keys_kernel<<<grid_for(total), 256, 0, st>>>(dist2, idx, n, k, drop, w.keys_a, w.vals_a);
  MGP_LAUNCH_CHECK();
  size_t cb = w.cub_bytes;
  // sort on all 64 bits so the all-ones sentinel (padding entries, idx = -1) lands at the end
  MGP_CUDA(cub::DeviceRadixSort::SortPairs(w.cub, cb, w.keys_a, w.keys_b, w.vals_a, w.vals_b, total, 0, 64, st));
  count_launch(4);
  sym_flag_heads_kernel<<<grid_for(total), 256, 0, st>>>(w.keys_b, total, w.flags);
  MGP_LAUNCH_CHECK();
  cb = w.cub_bytes;
  MGP_CUDA(cub::DeviceScan::ExclusiveSum(w.cub, cb, w.flags, w.pos, total, st));
  count_launch(2);
  sym_emit_kernel<<<grid_for(total), 256, 0, st>>>(w.keys_b, w.vals_b, w.flags, w.pos, total, n, cap, eidx, eval, m_out);
  MGP_LAUNCH_CHECK();
  return MGP_OK;
}

size_t mgp_csr_build_ws_bytes(int64_t n, int64_t m) {
  CsrWs w;
  if (n <= 0 || m <= 0) return 0;
  if (csr_layout(2 * m, nullptr, &w) != MGP_OK) return 0;
  return w.total_bytes;
}

int mgp_csr_build(const int64_t* eidx, int64_t ld, int64_t m, int64_t n, int32_t* rowptr, int32_t* col, int32_t* eid,
                  void* ws, size_t ws_bytes, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  MGP_CHECK_ARG(eidx && rowptr && col && eid && ws, "csr_build: null pointer");
  MGP_CHECK_ARG(n > 0 && m > 0 && ld >= m, "csr_build: bad sizes n=%lld m=%lld ld=%lld", (long long)n, (long long)m, (long long)ld);
  const int64_t nnz = 2 * m;
  MGP_CHECK_ARG(nnz < ((int64_t)1 << 31), "csr_build: 2M = %lld does not fit int32 row pointers", (long long)nnz);
  MGP_CHECK_ARG(n < ((int64_t)1 << 31), "csr_build: n must be < 2^31");
  CsrWs w;
  int rc = csr_layout(nnz, ws, &w);
  if (rc != MGP_OK) return rc;
  if (ws_bytes < w.total_bytes) { set_error("csr_build: workspace %zu < %zu", ws_bytes, w.total_bytes); return MGP_EWORKSPACE; }

  csr_make_keys_kernel<<<grid_for(m), 256, 0, st>>>(eidx, ld, m, n, w.keys_a, w.vals_a);
  MGP_LAUNCH_CHECK();
  size_t cb = w.cub_bytes;
  MGP_CUDA(cub::DeviceRadixSort::SortPairs(w.cub, cb, w.keys_a, w.keys_b, w.vals_a, eid, nnz, 0, key_bits(n), st));
  count_launch(4);
  csr_emit_kernel<<<grid_for(nnz), 256, 0, st>>>(w.keys_b, nnz, n, col);
  MGP_LAUNCH_CHECK();
  csr_rowptr_kernel<<<grid_for(n + 1), 256, 0, st>>>(w.keys_b, nnz, n, rowptr);
  MGP_LAUNCH_CHECK();
  return MGP_OK;
}

int mgp_gather_edge_f32(const float* val, const int32_t* eid, int64_t nnz, float* out, void* stream) {
  MGP_CHECK_ARG(val && eid && out && nnz > 0, "gather_edge: bad arguments");
  gather_edge_kernel<float><<<grid_for(nnz), 256, 0, (cudaStream_t)stream>>>(val, eid, nnz, out);
  MGP_LAUNCH_CHECK();
  return MGP_OK;
}

int mgp_gather_edge_f64(const double* val, const int32_t* eid, int64_t nnz, double* out, void* stream) {
  MGP_CHECK_ARG(val && eid && out && nnz > 0, "gather_edge: bad arguments");
  gather_edge_kernel<double><<<grid_for(nnz), 256, 0, (cudaStream_t)stream>>>(val, eid, nnz, out);
  MGP_LAUNCH_CHECK();
  return MGP_OK;
}

}  // extern "C"
