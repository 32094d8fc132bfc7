// Micro-benchmark: how fast can ONE persistent CTA per SM ingest a contiguous stream with cp.async.bulk (TMA, 1-D) into a
// shared-memory ring, as a function of chunk size, ring depth and the number of bulk ops per chunk?  Consumers only wait
// and release.  Also: the same stream with 16-byte cp.async (LDGSTS) from 128 producer threads, and a plain LDG.128 copy.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_stream tma_stream.cu ; run on the B200 box.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("{ .reg .b64 st; mbarrier.arrive.release.cta.shared::cta.b64 st, [%0]; }" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("{ .reg .b64 st; mbarrier.arrive.expect_tx.release.cta.shared::cta.b64 st, [%0], %1; }" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0;
  for (uint32_t spins = 0; !done; ++spins) {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.b32 %0, 1, 0, p; }" : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    if (!done && spins > (1u << 26)) __trap();
  }
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_arrive_noinc(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// mode 0: TMA, `pieces` bulk ops per chunk issued by one thread; mode 1: LDGSTS by 128 producer threads
template <int NCW>
__global__ void __launch_bounds__(128 + 32 * NCW, 1) ring_kernel(const unsigned char* src, size_t total, int chunk, int stages, int pieces, int mode,
                                                       unsigned long long* sink) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t full_bar[8], empty_bar[8];
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(&full_bar[s], mode == 0 ? 1 : 2 * 128); mbar_init(&empty_bar[s], NCW); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const size_t nchunks = total / chunk;
  unsigned long long acc = 0;
  if (warp < 4) {
    int it = 0;
    for (size_t c = blockIdx.x; c < nchunks; c += gridDim.x, ++it) {
      const int s = it % stages; const uint32_t ph = (it / stages) & 1;
      unsigned char* dst = smem + (size_t)s * chunk;
      const unsigned char* p = src + c * (size_t)chunk;
      mbar_wait(&empty_bar[s], ph ^ 1);
      if (mode == 0) {
        if (tid == 0) {
          mbar_arrive_expect_tx(&full_bar[s], (uint32_t)chunk);
          const int piece = chunk / pieces;
          for (int i = 0; i < pieces; ++i) bulk_g2s(dst + i * piece, p + i * piece, piece, &full_bar[s]);
        }
      } else if (mode == 1) {
        if (tid == 0) mbar_arrive_expect_tx(&full_bar[s], 0u); else mbar_arrive(&full_bar[s]);
        for (int o = tid * 16; o < chunk; o += 128 * 16) cp_async16(dst + o, p + o);
        cp_async_arrive_noinc(&full_bar[s]);
      } else {   // mode 2: TMA for `tma_part` of the chunk + scattered-looking cp.async for the rest, SpMM arrival protocol
        const int tma_bytes = (chunk * 3 / 4) & ~15;
        if (tid == 0) {
          mbar_arrive_expect_tx(&full_bar[s], (uint32_t)tma_bytes);
          const int piece = (tma_bytes / 3) & ~15;
          bulk_g2s(dst, p, piece, &full_bar[s]);
          bulk_g2s(dst + piece, p + piece, piece, &full_bar[s]);
          bulk_g2s(dst + 2 * piece, p + 2 * piece, tma_bytes - 2 * piece, &full_bar[s]);
        }
        for (int o = tma_bytes + tid * 16; o < chunk; o += 128 * 16) cp_async16(dst + o, p + o);
        cp_async_arrive_noinc(&full_bar[s]);
        if (tid != 0) mbar_arrive(&full_bar[s]);
      }
    }
  } else {
    int it = 0;
    for (size_t c = blockIdx.x; c < nchunks; c += gridDim.x, ++it) {
      const int s = it % stages; const uint32_t ph = (it / stages) & 1;
      mbar_wait(&full_bar[s], ph);
      acc += smem[(size_t)s * chunk + (tid & 31) * 4];
      __syncwarp();
      if ((tid & 31) == 0) mbar_arrive(&empty_bar[s]);
    }
  }
  if (acc == 0x12345678ull) *sink = acc;
}

__global__ void __launch_bounds__(1024, 2) ldg_kernel(const uint4* src, size_t n16, unsigned long long* sink) {
  unsigned long long acc = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
    const uint4 v = __ldcs(src + i);
    acc += v.x ^ v.y ^ v.z ^ v.w;
  }
  if (acc == 0x12345678ull) *sink = acc;
}

int main() {
  const size_t total = (size_t)1 << 30;   // 1 GiB stream (>> L2)
  unsigned char* src; unsigned long long* sink;
  cudaMalloc(&src, total); cudaMalloc(&sink, 8);
  cudaMemset(src, 1, total);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaFuncSetAttribute(ring_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  cudaFuncSetAttribute(ring_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  auto run = [&](const char* name, auto launch) {
    launch(); cudaDeviceSynchronize();
    cudaEventRecord(e0); for (int r = 0; r < 3; ++r) launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    cudaError_t err = cudaGetLastError();
    printf("%-44s %8.1f GB/s  %s\n", name, total * 3.0 / ms / 1e6, err == cudaSuccess ? "" : cudaGetErrorString(err));
  };
  char name[128];
  run("ldg.128 copy-read, 148x2x1024 threads", [&] { ldg_kernel<<<296, 1024>>>((const uint4*)src, total / 16, sink); });
  const int chunks[] = {32768};
  for (int ci = 0; ci < 1; ++ci)
    for (int stages = 3; stages <= 3; ++stages) {
      const int chunk = chunks[ci];
      if ((size_t)chunk * stages > 208 * 1024) continue;
      for (int pieces = 1; pieces <= 4; pieces *= 4) {
        snprintf(name, sizeof name, "tma  chunk=%5d stages=%d pieces=%d grid=148", chunk, stages, pieces);
        run(name, [&] { ring_kernel<1><<<148, 160, (size_t)chunk * stages>>>(src, total, chunk, stages, pieces, 0, sink); });
      }
      snprintf(name, sizeof name, "ldgsts chunk=%5d stages=%d grid=148", chunk, stages);
      run(name, [&] { ring_kernel<1><<<148, 160, (size_t)chunk * stages>>>(src, total, chunk, stages, 1, 1, sink); });
    }
  for (int chunk : {32768, 49152, 65536})
    for (int mode = 0; mode <= 2; ++mode) {
      snprintf(name, sizeof name, "16 consumer warps: mode=%d chunk=%5d stages=3", mode, chunk);
      run(name, [&] { ring_kernel<16><<<148, 128 + 512, (size_t)chunk * 3>>>(src, total, chunk, 3, 1, mode, sink); });
    }
  return 0;
}
