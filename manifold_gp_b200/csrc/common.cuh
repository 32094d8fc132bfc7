// Shared device/host helpers for libmgp_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/mgp_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libmgp_b200 targets sm_100a only"
#endif

namespace mgp {

// ---- error plumbing ------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define MGP_CHECK_ARG(cond, ...)            \
  do {                                      \
    if (!(cond)) {                          \
      ::mgp::set_error(__VA_ARGS__);        \
      return MGP_EINVAL;                    \
    }                                       \
  } while (0)

#define MGP_CUDA(call)                                                                         \
  do {                                                                                         \
    cudaError_t e__ = (call);                                                                  \
    if (e__ != cudaSuccess) {                                                                  \
      ::mgp::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
      return MGP_ECUDA;                                                                        \
    }                                                                                          \
  } while (0)

// after a kernel launch
#define MGP_LAUNCH_CHECK()                 \
  do {                                     \
    ::mgp::count_launch();                 \
    MGP_CUDA(cudaGetLastError());          \
  } while (0)

constexpr int kNumSMs = 148;  // B200

// ---- programmatic dependent launch (PDL) ------------------------------------------------------------------------------
// The solver loops are chains of 3-4 dependent launches of 20-130 us each; launched with the programmatic-stream-serialization
// attribute a kernel's CTAs are scheduled while its predecessor drains, and block in `griddepcontrol.wait` (first statement of the
// kernel) until the predecessor has completed and flushed -- the launch latency between dependent kernels disappears from the
// critical path.  MGP_PDL=0 restores plain launches.  A kernel launched this way MUST call pdl_wait() before touching anything
// a predecessor wrote; pdl_launch_dependents() lets ITS successor's CTAs become resident early (they block in their own wait).
bool pdl_enabled();
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- device helpers ------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// reduce across lanes that differ in the bits [lo, hi) of the lane id (both powers of two, lo < hi <= 32)
template <typename T>
__device__ __forceinline__ T subwarp_sum(T v, int lo, int hi) {
  for (int o = lo; o < hi; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// streaming (read-once) loads: bypass L1 allocation so the gather working set keeps the cache
__device__ __forceinline__ int ld_stream(const int* p) {
  int v;
  asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ float ld_stream(const float* p) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ double ld_stream(const double* p) {
  double v;
  asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ int4 ld_stream_v4(const int4* p) {
  int4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

template <typename T>
__device__ __forceinline__ T dev_exp(T x);
template <>
__device__ __forceinline__ float dev_exp<float>(float x) { return expf(x); }
template <>
__device__ __forceinline__ double dev_exp<double>(double x) { return exp(x); }
template <typename T>
__device__ __forceinline__ T dev_sqrt(T x);
template <>
__device__ __forceinline__ float dev_sqrt<float>(float x) { return sqrtf(x); }
template <>
__device__ __forceinline__ double dev_sqrt<double>(double x) { return sqrt(x); }

// Deterministic grid-wide reduction helper ("last block" pattern).
// Each block writes `ncols` partials to partials[blockIdx.x * ncols + c], then calls last_block_ticket();
// exactly one block (the last to arrive) gets `true` and may reduce partials[0 .. gridDim.x) in fixed order.
// `counter` must be zero on entry and is reset to zero by the last block.
__device__ __forceinline__ bool last_block_ticket(unsigned int* counter) {
  __shared__ bool is_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int t = atomicAdd(counter, 1u);
    is_last = (t == gridDim.x - 1);
    if (is_last) *counter = 0u;
  }
  __syncthreads();
  if (is_last) __threadfence();
  return is_last;
}

// Same ticket when only SOME threads wrote the partials (`wrote`): only they fence.  A fence waits for ALL outstanding stores of
// the calling thread, so in the SpMM kernels -- where the partials are written by 16 threads of a producer warp while the 512
// consumer threads still have their Y rows in flight -- fencing everybody put the drain of the whole tile's output on the
// critical path of every block (measured at 125k rows per rank: 3 us per launch for the dot epilogue).
__device__ __forceinline__ bool last_block_ticket_writers(unsigned int* counter, bool wrote) {
  __shared__ bool is_last_w;
  if (wrote) __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int t = atomicAdd(counter, 1u);
    is_last_w = (t == gridDim.x - 1);
    if (is_last_w) *counter = 0u;
  }
  __syncthreads();
  if (is_last_w) __threadfence();
  return is_last_w;
}

// system-scope release / acquire on flags in peer-mapped memory (multi-GPU sync points)
__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

}  // namespace mgp
